#!/bin/bash
# Long fuzz sweep (GPU): N extra random triangle / segment soups through every test of tests/test_gpu_fuzz.py.
#   usage: tools/gpu_fuzz_sweep.sh <tag> [N] [first seed]
TAG=${1:-fuzz}; N=${2:-150}; export SNCH_FUZZ_FIRST=${3:-101}
mkdir -p gpurun_out/$TAG
SNCH_FUZZ_EXTRA=$N timeout 1500 python -m pytest tests/test_gpu_fuzz.py -m gpu -q --timeout 1200 > gpurun_out/$TAG/fuzz_sweep.log 2>&1
echo "exit $?" >> gpurun_out/$TAG/fuzz_sweep.log
tail -25 gpurun_out/$TAG/fuzz_sweep.log
