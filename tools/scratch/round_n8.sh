set -u
OUT=gpurun_out/r01u_n8
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 --no-extra > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "bench n8 exit $?"; tail -3 $OUT/bench_n8.err | cut -c1-300
python -c "
import json;d=json.load(open('$OUT/bench_n8.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e'],d['extra'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > $OUT/bench_ref_n8.json 2> $OUT/bench_ref_n8.err; echo "ref n8 exit $?"; tail -c 300 $OUT/bench_ref_n8.json
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 --no-extra --queries 4194304 2>&1 | grep -i "NVLS\|nvlink\|via P2P\|Channel 00/" | head -8 > $OUT/nccl_info.txt; cat $OUT/nccl_info.txt | cut -c1-200
