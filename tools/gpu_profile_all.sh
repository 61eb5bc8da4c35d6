#!/bin/bash
# One gpurun call: `ncu --set full` of every kernel family outside the two headline kernels (tools/profile_all.py), condensed on
# the box into raw CSVs (the .ncu-rep files travel too when small enough).   usage: tools/gpu_profile_all.sh <tag>
set -u
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1
timeout 900 ncu --set full --profile-from-start off --clock-control none --import-source on -k regex:k_intersect -o "$OUT/prof_ray" \
  python tools/profile_all.py --groups ray,ray4m > "$OUT/prof_ray.log" 2>&1
echo "ncu ray exit $?"
timeout 1500 ncu --set full --profile-from-start off --clock-control none -o "$OUT/prof_rest" \
  python tools/profile_all.py --groups sample,wide,build,adjacency,2d > "$OUT/prof_rest.log" 2>&1
echo "ncu rest exit $?"
for r in prof_ray prof_rest; do
  ncu -i "$OUT/$r.ncu-rep" --page raw --csv > "$OUT/$r.raw.csv" 2>/dev/null
done
python tools/ncu_lines.py "$OUT/prof_ray.ncu-rep" snch-lbvh_b200/csrc/query.o k_intersect --top 50 --json "$OUT/ray_lines.json" > "$OUT/ray_lines.txt" 2>&1
rm -f "$OUT/prof_rest.ncu-rep"   # 90 MB for ~60 kernels: only the raw CSV travels (gpurun_out is capped at 64 MiB)
[ "${KEEP_REP:-0}" = "1" ] || rm -f "$OUT/prof_ray.ncu-rep"
ls -la "$OUT"
