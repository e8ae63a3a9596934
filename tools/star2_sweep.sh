#!/bin/bash
# (row warps, ring stages, quads per lane) sweep of the fused star-CTC kernels on BASELINE config 3; `ncu` as 2nd argument adds one
# full ncu capture.  Usage (under gpurun): bash tools/star2_sweep.sh [tag] [ncu] [configs...]
O=gpurun_out; T=${1:-a}; NCU=$2; shift; shift
export PYTHONPATH=.
CFGS=("$@"); [ ${#CFGS[@]} -eq 0 ] && CFGS=("8 3 4" "8 3 2" "8 3 1" "4 3 2")
for cfg in "${CFGS[@]}"; do
  set -- $cfg
  HA_B200_STAR_R=$1 HA_B200_STAR_NS=$2 HA_B200_STAR_J=$3 timeout 120 python bench.py --workload star --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-library-baseline \
      > $O/star2_R$1_NS$2_J$3_$T.json 2> $O/star2_R$1_NS$2_J$3_$T.err
  python - "$O/star2_R$1_NS$2_J$3_$T.json" "$cfg" <<'PY'
import json, sys
try:
    d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print("R NS J =", sys.argv[2], "ms/step %.4f" % d["ms_per_step"], "fwd %.4f bwd %.4f" % (d.get("fwd_ms", -1), d.get("bwd_ms", -1)), "frac %.3f" % d["roofline"]["frac"])
except Exception as e:
    print("R NS J =", sys.argv[2], "failed:", e)
PY
done
if [ "$NCU" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'star2_(fwd|bwd)' -s 4 -c 2 -o $O/prof_star2_$T \
    python bench.py --workload star --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-library-baseline > $O/ncu_star2_$T.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $O/prof_star2_$T.ncu-rep > $O/ncu_star2_$T.txt 2>/dev/null
python tools/ncu_lines.py $O/prof_star2_$T.ncu-rep "" 60 > $O/ncu_star2_lines_$T.txt 2>/dev/null
fi
