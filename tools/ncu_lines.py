"""Per-source-line instruction and stall-sample totals of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py rep.ncu-rep [kernel-substring] [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fn = None; path = None; hdr = None
agg = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, "", collections.Counter()]))
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path": path = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": fn = row[1]; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or len(row) < len(hdr) - 2: continue
    d = dict(zip(hdr, row))
    # two "Source" columns: first is CUDA source, second SASS; dict keeps the last -> use indices
    line = row[0]; src = row[1]
    if not line.strip(): continue          # SASS rows under a source line: the line's own row carries the totals
    try:
        ie = int(row[hdr.index("Instructions Executed")]); ns = int(row[hdr.index("# Samples")])
    except ValueError:
        continue
    a = agg[fn][(path, line)]
    a[0] += ie; a[1] += ns; a[2] = src
    for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_membar", "stall_sleep", "stall_branch_resolving", "stall_not_selected", "stall_dispatch"):
        try: a[3][k] += int(row[hdr.index(k)])
        except (ValueError, IndexError): pass
for fn, lines in agg.items():
    if pat not in fn: continue
    ti = sum(a[0] for a in lines.values()); ts = sum(a[1] for a in lines.values())
    print(f"== {fn}: {ti} warp instructions, {ts} samples")
    for (path, line), a in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        st = ", ".join(f"{k[6:]}={v}" for k, v in a[3].most_common(3) if v)
        print(f"{path}:{line:>4s} inst {100*a[0]/max(ti,1):5.1f}%  samp {100*a[1]/max(ts,1):5.1f}%  [{st}]  {a[2].strip()[:90]}")
