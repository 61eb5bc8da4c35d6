#!/bin/bash
# One gpurun call: GPU tests, the bench line, the ncu launch list of the same command, one `--set full` capture of the
# dominant kernel.  Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ by tools/ncu_summary.py.
#   usage: tools/gpu_round.sh <tag> [skip-tests]
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1
if [ "${2:-}" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -5 "$OUT/pytest_gpu.log"
fi
timeout 900 python bench.py --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "bench exit $?"; tail -c 3000 "$OUT/bench.json"; tail -5 "$OUT/bench.err"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
echo "bench ref exit $?"; tail -c 1500 "$OUT/bench_reference.json"
# launch list of the same command (shares only; a number printed under ncu is never a bench value)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
  python bench.py --steps 3 --warmup 3 --no-extra > "$OUT/bench_under_ncu.log" 2>&1
echo "ncu launches exit $?"
# full capture of the dominant kernel (3rd launch = after warm-up)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_silhouette -s 2 -c 1 -o "$OUT/prof_silhouette" \
  python bench.py --steps 1 --warmup 3 --no-extra > "$OUT/ncu_full.log" 2>&1
echo "ncu full exit $?"
ls -la "$OUT"
