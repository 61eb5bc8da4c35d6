set -u
OUT=gpurun_out/r01l
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --also c4,c5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['value'],d['e2e']);print(json.dumps(d['extra'].get('c4')));print(json.dumps(d['extra'].get('c5')))"
