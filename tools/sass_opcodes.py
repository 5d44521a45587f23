"""Opcode histogram of the Blackwell-specific / notable SASS instructions per kernel of the built library.
usage: python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt   (needs cuobjdump; runs without a GPU)"""
import collections, os, re, subprocess, sys
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(here, "ddpm_torch_b200", "libddpm_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
NOTABLE = ("UTC", "UTMA", "LDTM", "STTM", "UCGABAR", "ACQBULK", "PREEXIT", "RED", "ATOMG", "SHFL.BFLY", "MUFU.TANH", "UBLKCP", "SYNCS", "LDG.E.ENL2.256", "STG.E.ENL2.256")
print("cuobjdump -sass ddpm_torch_b200/libddpm_b200.so : Blackwell-specific / notable opcodes per kernel (build of this commit)")
print("UTCHMMA = tcgen05.mma, .2CTA = cta_group::2; UTMALDG/UTMASTG = TMA load/store; LDTM = tcgen05.ld; UTCBAR = tcgen05.commit; UTCATOMSWS = TMEM alloc; UCGABAR = cluster barrier\n")
cur = None; counts = collections.OrderedDict()
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        if op.startswith(NOTABLE): counts[cur][op] += 1
for fn, c in counts.items():
    if not c: continue
    name = demangle(fn)
    name = re.sub(r"\(.*", "", name)
    print(name)
    print("   " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
