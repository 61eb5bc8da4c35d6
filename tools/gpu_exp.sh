#!/bin/bash
# One gpurun call for kernel experiments: the knob-sweep parity test, the A/B table of the named variant sets, and a
# `--set full` capture of the silhouette kernel condensed ON THE BOX (details / raw CSV + per-source-line table), so
# the large .ncu-rep need not travel.
#   usage: tools/gpu_exp.sh <tag> <variant sets> [ncu knob=value ...]
set -u
TAG=${1:-exp}
SETS=${2:-default}
shift 2
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1
timeout 900 python -m pytest tests/test_gpu_queries.py -m gpu -q -x > "$OUT/pytest_queries.log" 2>&1
echo "pytest exit $?"; tail -5 "$OUT/pytest_queries.log"
timeout 900 python tools/variants.py --sets "$SETS" > "$OUT/variants.json" 2> "$OUT/variants.err"
echo "variants exit $?"; cut -c1-400 "$OUT/variants.err"
if [ "${NCU:-1}" = "1" ]; then
  SNCH_OPTIONS="$*" timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_silhouette -s 2 -c 1 -o "$OUT/prof_silhouette" \
    python bench.py --steps 1 --warmup 3 --no-extra > "$OUT/ncu_full.log" 2>&1
  echo "ncu full silhouette exit $?"
  ncu -i "$OUT/prof_silhouette.ncu-rep" --page details --csv > "$OUT/sil_details.csv" 2>/dev/null
  ncu -i "$OUT/prof_silhouette.ncu-rep" --page raw --csv > "$OUT/sil_raw.csv" 2>/dev/null
  python tools/ncu_lines.py "$OUT/prof_silhouette.ncu-rep" snch-lbvh_b200/csrc/query.o "${NCU_KERNEL:-k_silhouette_coopILi2ELb0ELb0}" --top 60 --json "$OUT/sil_lines.json" > "$OUT/sil_lines.txt" 2>&1
  ls -la "$OUT"
  [ "${KEEP_REP:-0}" = "1" ] || rm -f "$OUT/prof_silhouette.ncu-rep"
fi
