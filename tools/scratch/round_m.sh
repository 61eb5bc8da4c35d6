set -u
OUT=gpurun_out/r01m
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_build.py tests/test_gpu_refit_serialise.py tests/test_gpu_cpp_dropin.py -m gpu -q -x > $OUT/pytest_build.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_build.log; tail -5 $OUT/pytest_build.log
python tools/scratch/build_only.py > $OUT/build_only.log 2>&1; cat $OUT/build_only.log
python tools/scratch/build_only.py 2240 > $OUT/build_only_10m.log 2>&1; cat $OUT/build_only_10m.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_refit -s 2 -c 1 -o $OUT/prof_refit python tools/scratch/build_only.py > $OUT/ncu_refit.log 2>&1; echo "ncu refit $?"
timeout 600 python tools/scratch/e2e_sweep.py > $OUT/e2e_sweep.log 2>&1; cat $OUT/e2e_sweep.log
