#!/bin/bash
# Round-end measurement pass on the GPU box: benches, launch lists, full ncu captures.
# Usage (under gpurun): bash tools/round_profile.sh r01
R=${1:-r01}
O=gpurun_out
for w in ctc star rnnt; do
  timeout 500 python bench.py --workload $w --steps 100 --warmup 5 > $O/bench_${w}_$R.json 2> $O/bench_${w}_$R.err; echo "bench $w rc=$?"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_$R.json 2> $O/bench_reference_$R.err; echo "reference rc=$?"
for w in ctc star rnnt; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_${w}_$R.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1; echo "launches $w rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ctc_(grad|rows|trellis)' -s 4 -c 3 -o $O/prof_ctc_$R \
    python bench.py --workload ctc --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > $O/ncu_ctc_$R.log 2>&1; echo "ncu ctc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rnnt_(grad|rows|lattice)' -s 4 -c 3 -o $O/prof_rnnt_$R \
    python bench.py --workload rnnt --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > $O/ncu_rnnt_$R.log 2>&1; echo "ncu rnnt rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'star_(grad|rows|trellis)' -s 4 -c 3 -o $O/prof_star_$R \
    python bench.py --workload star --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > $O/ncu_star_$R.log 2>&1; echo "ncu star rc=$?"
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu_$R.csv
