"""SASS census of the shipped library (cuobjdump -sass): per-kernel counts of the mnemonics that prove which hardware
paths are used — UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UTMALDG (tensor-map TMA),
UBLKCP (bulk TMA, cp.async.bulk), SYNCS (mbarrier), CREDUX, MUFU, FFMA.  usage: python tools/sass_census.py > profiles/sass_census_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "haloop_b200", "libha_b200.so")
KEYS = ("CREDUX", "FFMA", "LDTM", "MUFU", "STTM", "SYNCS", "UBLKCP", "UTCBAR", "UTCHMMA", "UTMALDG")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
per = collections.OrderedDict(); cur = None; k = 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names[k]; k += 1; per[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        for key in KEYS:
            if op.startswith(key):
                per[cur][key] += 1
tot = collections.Counter()
for c in per.values():
    tot.update(c)
fmt = lambda c: ", ".join(f"{key} {c[key]}" for key in KEYS if c[key])
print("# SASS census of haloop_b200/libha_b200.so (cuobjdump -sass, sm_100a), round 2 — tools/sass_census.py\n#")
print("# whole library:", fmt(tot), "\n#")
print("# kernels that use the tensor cores / tensor memory / tensor-map TMA (tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM,\n"
      "# cp.async.bulk.tensor = UTMALDG, tcgen05.commit = UTCBAR):")
for n, c in per.items():
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
        print(n); print("   ", fmt(c))
print("#\n# kernels that stream rows with bulk TMA copies (cp.async.bulk = UBLKCP) and mbarriers (SYNCS):")
for n, c in per.items():
    if c["UBLKCP"] and not (c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]):
        print(n); print("   ", fmt(c))
