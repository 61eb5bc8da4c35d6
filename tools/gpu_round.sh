#!/bin/bash
# One gpurun call: GPU tests, the bench line, the ncu launch list of the same command, `--set full` captures of the
# dominant kernels, and the knob A/B table.  Everything lands in gpurun_out/<tag>/ (scratch); tools/ncu_summary.py
# condenses it into profiles/.
#   usage: tools/gpu_round.sh <tag> [steps: tests,bench,ref,launches,ncu,variants (default all)]
set -u
TAG=${1:-run}
STEPS=${2:-tests,bench,ref,launches,ncu,variants}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1
has() { case ",$STEPS," in *",$1,"*) return 0;; *) return 1;; esac; }
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -15 "$OUT/pytest_gpu.log"
fi
if has variants; then
  timeout 900 python tools/variants.py > "$OUT/variants.json" 2> "$OUT/variants.err"
  echo "variants exit $?"; cat "$OUT/variants.err" | cut -c1-600
fi
if has bench; then
  timeout 1200 python bench.py --steps 10 --warmup 3 --also 2d > "$OUT/bench.json" 2> "$OUT/bench.err"
  echo "bench exit $?"; tail -c 3500 "$OUT/bench.json"; tail -5 "$OUT/bench.err"
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
  echo "bench ref exit $?"; tail -c 1500 "$OUT/bench_reference.json"
fi
if has launches; then
  # launch list of the same command (shares only; a number printed under ncu is never a bench value)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-extra > "$OUT/bench_under_ncu.log" 2>&1
  echo "ncu launches exit $?"
fi
if has ncu; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_silhouette -s 2 -c 1 -o "$OUT/prof_silhouette" \
    python bench.py --steps 1 --warmup 3 --no-extra > "$OUT/ncu_full.log" 2>&1
  echo "ncu full silhouette exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_closest -c 1 -o "$OUT/prof_closest" \
    python bench.py --steps 1 --warmup 3 --no-extra > "$OUT/ncu_full_closest.log" 2>&1
  echo "ncu full closest exit $?"
fi
ls -la "$OUT"
if has parity; then
  timeout 900 python tools/parity_report.py > "$OUT/parity_report.log" 2>&1
  echo "parity report exit $?"; cp gpurun_out/parity_report.json "$OUT/parity_report.json" 2>/dev/null
fi
