#!/bin/bash
# Round-end measurement pass on the GPU box: benches, launch lists, full ncu captures.
# Usage (under gpurun): bash tools/round_profile.sh r01      then, here:  bash tools/round_profile.sh collect r01
if [ "$1" = "collect" ]; then
  R=${2:-r01}
  for w in ctc star rnnt rnnt_fg head; do
    [ -s gpurun_out/ncu_${w}_$R.txt ] && cp gpurun_out/ncu_${w}_$R.txt profiles/
    [ -f gpurun_out/launches_${w}_$R.csv ] && cp gpurun_out/launches_${w}_$R.csv profiles/
  done
  [ -s gpurun_out/bench_default_$R.json ] && cp gpurun_out/bench_default_$R.json profiles/
  for w in ctc star rnnt rnnt_fg ctc_c1_graph sweep_ctc sweep_rnnt reference; do
    [ -s gpurun_out/bench_${w}_$R.json ] && cp gpurun_out/bench_${w}_$R.json profiles/
  done
  cp gpurun_out/gpu_$R.csv profiles/ 2>/dev/null
  python tools/make_traffic.py $R > /dev/null
  exit 0
fi
R=${1:-r01}
O=gpurun_out
export PYTHONPATH=.
if [ "$2" = "star" ]; then      # refresh of the star-CTC artefacts only (+ the default line, which carries the star sub-record)
  timeout 900 python bench.py > $O/bench_default_$R.json 2> $O/bench_default_$R.err; echo "bench default rc=$?"
  timeout 500 python bench.py --workload star --steps 100 --warmup 5 > $O/bench_star_$R.json 2> $O/bench_star_$R.err; echo "bench star rc=$?"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_star_$R.csv \
      python bench.py --workload star --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-library-baseline > /dev/null 2>&1; echo "launches star rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'star2_(fwd|bwd)' -s 4 -c 2 -o $O/prof_star_$R \
      python bench.py --workload star --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-library-baseline > $O/ncu_star_$R.log 2>&1; echo "ncu star rc=$?"
  python tools/ncu_summary.py $O/prof_star_$R.ncu-rep > $O/ncu_star_$R.txt 2>/dev/null
  python tools/ncu_lines.py $O/prof_star_$R.ncu-rep "" 40 > $O/ncu_star_lines_$R.txt 2>/dev/null; rm -f $O/prof_star_$R.ncu-rep
  for tool in memcheck racecheck synccheck; do
    timeout 600 compute-sanitizer --tool $tool python tools/sanitize_cases.py star > $O/san_${tool}_star_$R.log 2>&1; echo "$tool star rc=$? $(grep -c "ERROR SUMMARY" $O/san_${tool}_star_$R.log)"
  done
  exit 0
fi
timeout 900 python bench.py > $O/bench_default_$R.json 2> $O/bench_default_$R.err; echo "bench default (headline + workloads) rc=$?"
for w in ctc star rnnt rnnt_fg; do
  timeout 500 python bench.py --workload $w --steps 100 --warmup 5 > $O/bench_${w}_$R.json 2> $O/bench_${w}_$R.err; echo "bench $w rc=$?"
done
timeout 300 python bench.py --workload ctc_c1 --graph --steps 200 --warmup 10 --no-e2e --no-cpu-baseline --no-library-baseline > $O/bench_ctc_c1_graph_$R.json 2> $O/bench_ctc_c1_graph_$R.err; echo "bench ctc_c1 graph rc=$?"
for w in sweep_ctc sweep_rnnt; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 2 > $O/bench_${w}_$R.json 2> $O/bench_${w}_$R.err; echo "bench $w rc=$?"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_$R.json 2> $O/bench_reference_$R.err; echo "reference rc=$?"
for w in ctc star rnnt rnnt_fg; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_${w}_$R.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-library-baseline > /dev/null 2>&1; echo "launches $w rc=$?"
done
timeout 600 ncu --set full --clock-control none -k regex:'ctc2_(fwd|bwd)' -s 4 -c 2 -o $O/prof_ctc_$R \
    python bench.py --workload ctc --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-library-baseline > $O/ncu_ctc_$R.log 2>&1; echo "ncu ctc rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'rnnt_(grad|rows|lattice)' -s 4 -c 3 -o $O/prof_rnnt_$R \
    python bench.py --workload rnnt --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-library-baseline > $O/ncu_rnnt_$R.log 2>&1; echo "ncu rnnt rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'star2_(fwd|bwd)' -s 4 -c 2 -o $O/prof_star_$R \
    python bench.py --workload star --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-library-baseline > $O/ncu_star_$R.log 2>&1; echo "ncu star rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'umma_gemm|fg_rows|fg_arc|fg_w_' -s 12 -c 6 -o $O/prof_rnnt_fg_$R \
    python bench.py --workload rnnt_fg --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-library-baseline > $O/ncu_rnnt_fg_$R.log 2>&1; echo "ncu rnnt_fg rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'umma_gemm|head_|ctc_' --csv --log-file $O/launches_head_$R.csv \
    python tools/head_check.py perf > /dev/null 2>&1; echo "launches head rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'umma_gemm' -c 6 -o $O/prof_head_$R \
    python tools/head_check.py prof > $O/ncu_head_$R.log 2>&1; echo "ncu head rc=$?"
# the reports are summarised on the box (gpurun brings back at most 64 MiB) and dropped
for w in ctc star rnnt rnnt_fg head; do
  python tools/ncu_summary.py $O/prof_${w}_$R.ncu-rep > $O/ncu_${w}_$R.txt 2>/dev/null; rm -f $O/prof_${w}_$R.ncu-rep
done
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu_$R.csv
