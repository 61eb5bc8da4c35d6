set -u
OUT=gpurun_out/r01q
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra > $OUT/bench_noextra.json 2> $OUT/bench.err; tail -3 $OUT/bench.err; python -c "
import json;d=json.load(open('$OUT/bench_noextra.json'));print(d['value'],d['ms_per_step'],d['e2e'],d['gpu_launches'],d['roofline']['kernel_ms'])"
