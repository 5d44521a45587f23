import csv, collections, re, sys
f=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 22
lines=[l for l in open(f) if not l.startswith("==")]
rows=[]
for r in csv.DictReader(lines):
    if r.get("Metric Name")=="gpu__time_duration.sum":
        v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
        v = v/1e3 if u=="ns" else (v*1e3 if u=="ms" else v)
        rows.append((r["Kernel Name"], v, r["Grid Size"]))
tot=sum(v for _,v,_ in rows)
agg=collections.defaultdict(lambda:[0,0.0])
for k,v,_ in rows:
    k2=re.sub(r"\(.*","",k)[:60]; agg[k2][0]+=1; agg[k2][1]+=v
print("launches",len(rows),"total us %.0f"%tot)
for k,(n,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:top]:
    print("  %-62s n=%4d  %8.0f us  %5.1f%%"%(k,n,v,100*v/tot))
