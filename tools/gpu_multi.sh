#!/bin/bash
# One gpurun --gpus N call: the strong-scaling bench at N ranks (torchrun, as the driver launches it), the two-process
# replication test, and NCCL's own transport report.   usage: tools/gpu_multi.sh <tag> <N> [steps]
set -u
TAG=${1:-multi}
N=${2:-2}
STEPS=${3:-5}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > "$OUT/smi.txt" 2>&1
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
timeout 600 python -m pytest tests/test_gpu_replication.py -m gpu -q > "$OUT/pytest_replication.log" 2>&1
echo "pytest exit $?"; tail -5 "$OUT/pytest_replication.log"
NCCL_DEBUG=INFO NCCL_DEBUG_FILE="$OUT/nccl_%h_%p.txt" timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus "$N" --steps "$STEPS" --warmup 3 ${BENCH_FLAGS:-} > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"
echo "bench N=$N exit $?"; tail -c 6000 "$OUT/bench_n$N.json"; tail -15 "$OUT/bench_n$N.err"
grep -h -E "NVLS|via P2P|via SHM|Channel 00|Connected all|nranks|comm 0x" "$OUT"/nccl_*.txt 2>/dev/null | sort | uniq -c | sort -rn | head -30 > "$OUT/nccl_summary.txt"
rm -f "$OUT"/nccl_*_*.txt
ls -la "$OUT"
