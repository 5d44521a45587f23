"""Main-stream timeline of one training step with the side stream live: DDPM_OP_TIMING event pairs per op,
averaged over a few steps and aggregated by op kind.  usage: python tools/op_timing.py [train|fwd] [B]"""
import os, sys, re, collections
path = "/tmp/ddpm_op_timing.txt"
os.environ["DDPM_OP_TIMING"] = path
sys.argv = [sys.argv[0]] + (sys.argv[1:] if len(sys.argv) > 1 else ["train", "128"])
here = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(here, "profile_step.py")).read().split("for i in range(2):")[0]
exec(src)
for i in range(3):
    step(i)
torch.cuda.synchronize()
open(path, "w").close()
N = 5
for i in range(N):
    step(10 + i)
torch.cuda.synchronize()
tot = collections.OrderedDict(); kinds = collections.Counter(); kcount = collections.Counter()
for ln in open(path):
    name, ms = ln.rsplit(" ", 1)
    tot[name] = tot.get(name, 0.0) + float(ms) / N
def kind(n):
    if n.startswith("[side]"): return "[side] (issue only)"
    if n.startswith("[join]"): return "[join wait]"
    parts = n.split(".")
    k = parts[-1]
    if k in ("bwd", "fwd") and len(parts) > 1: k = parts[-2] + "." + k
    return k
for n, v in tot.items():
    kinds[kind(n)] += v; kcount[kind(n)] += 1
print("total main-stream ms/step: %.3f" % sum(tot.values()))
for k, v in kinds.most_common():
    print("%-28s %4d ops  %8.3f ms" % (k, kcount[k], v))
print("---- top 40 ops")
for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print("%-60s %8.4f" % (n, v))
out = os.path.join(os.path.dirname(here), "gpurun_out", "op_timing_%s.txt" % mode)
with open(out, "w") as f:
    for n, v in tot.items(): f.write("%s %.5f\n" % (n, v))
