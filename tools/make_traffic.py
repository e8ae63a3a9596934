"""profiles/ncu_{ctc,star,rnnt}_rNN.txt (tools/ncu_summary.py output) -> profiles/traffic_rNN.json:
DRAM bytes per launch of every profiled kernel, read by bench.py for roofline.traffic.
usage: python tools/make_traffic.py r01"""
import json, os, re, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
out = {}
for w in ("ctc", "star", "rnnt", "rnnt_fg", "head"):
    path = os.path.join(root, f"ncu_{w}_{R}.txt")
    if not os.path.exists(path):
        continue
    name = None
    for line in open(path):
        m = re.match(r"kernel\s+(?:void\s+)?(\w+)", line)
        if m:
            name = m.group(1)
        m = re.match(r"dram_total\s+([0-9.]+) GB", line)
        if m and name:
            out[name] = {"dram_bytes_per_launch": float(m.group(1)) * 1e9,
                         "source": f"profiles/ncu_{w}_{R}.txt (ncu --set full --clock-control none, BASELINE config, one launch)"}
json.dump(out, open(os.path.join(root, f"traffic_{R}.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
