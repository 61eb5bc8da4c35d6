#!/bin/bash
# The evidence run of a round on ONE GPU: tests, the bench line and its CPU arm, the ncu launch list of the same command,
# `--set full` captures of the three headline kernels (+ every other kernel family as raw CSV), the parity report, and the
# small-shard knob table.   usage: tools/gpu_final.sh <tag>
set -u
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
bash tools/gpu_round.sh "$TAG" tests,bench,ref,launches,ncu
SNCH_NCU_QUERIES=16777216 timeout 900 ncu --set full --profile-from-start off --clock-control none --import-source on -k regex:k_intersect -o "$OUT/prof_ray" \
  python tools/profile_all.py --groups ray,ray4m > "$OUT/prof_ray.log" 2>&1
echo "ncu ray exit $?"
timeout 1500 ncu --set full --profile-from-start off --clock-control none -o "$OUT/prof_rest" \
  python tools/profile_all.py --groups sample,wide,build,adjacency,2d > "$OUT/prof_rest.log" 2>&1
echo "ncu rest exit $?"
ncu -i "$OUT/prof_rest.ncu-rep" --page raw --csv > "$OUT/prof_rest.raw.csv" 2>/dev/null
python tools/ncu_lines.py "$OUT/prof_ray.ncu-rep" snch-lbvh_b200/csrc/query.o k_intersect_parked --top 40 > "$OUT/ray_lines.txt" 2>&1
python tools/ncu_lines.py "$OUT/prof_silhouette.ncu-rep" snch-lbvh_b200/csrc/query.o k_silhouette_coop --top 50 > "$OUT/sil_lines.txt" 2>&1
rm -f "$OUT/prof_rest.ncu-rep"
timeout 1200 python tools/parity_report.py --big > "$OUT/parity_report.log" 2>&1
echo "parity report exit $?"; cp gpurun_out/parity_report.json "$OUT/parity_report.json" 2>/dev/null
timeout 600 python tools/variants.py --queries 2097152 --sets default,sil_tail0,sil_tail8,sil_tail31,sil_flush16,sil_flush32,ray_v1 > "$OUT/variants_2m.json" 2> "$OUT/variants_2m.err"
tail -8 "$OUT/variants_2m.err" | cut -c1-300
du -sh "$OUT"; ls -la "$OUT"
