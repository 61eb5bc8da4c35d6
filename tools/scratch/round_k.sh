set -u
OUT=gpurun_out/r01k
mkdir -p $OUT
python tools/scratch/adj_timing.py > $OUT/adj_timing.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
python tools/scratch/build_only.py > $OUT/build_only.log 2>&1; cat $OUT/build_only.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_refit -s 2 -c 1 -o $OUT/prof_refit python tools/scratch/build_only.py > $OUT/ncu_refit.log 2>&1; echo "ncu refit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_hierarchy|k_radix_scatter|k_radix_hist|k_morton|k_scene_box" -s 12 -c 12 -o $OUT/prof_build_misc python tools/scratch/build_only.py > $OUT/ncu_misc.log 2>&1; echo "ncu misc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/build_launches.csv python tools/scratch/build_only.py > /dev/null 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['value'],d['e2e'],d['extra'].get('build_ms'))"
