set -u
OUT=gpurun_out/r01p
mkdir -p $OUT
timeout 300 python tools/scratch/build_only.py > $OUT/build_only.log 2>&1; cat $OUT/build_only.log
timeout 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_adjacency.py tests/test_gpu_refit_serialise.py tests/test_gpu_cpp_dropin.py tests/test_gpu_reference_cuda.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
timeout 300 python tools/scratch/build_only.py 2240 > $OUT/build_only_10m.log 2>&1; tail -2 $OUT/build_only_10m.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/build_launches.csv python tools/scratch/build_only.py > /dev/null 2>&1
