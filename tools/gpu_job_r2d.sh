#!/bin/bash
# round-2 GPU job D (2 GPUs): bench.py as the driver launches it at N=2 (configs[3] sharded + sub-records), and the
# multi-GPU correctness tool
O=gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 20 --warmup 5 > $O/r2d_bench_${N}gpu.json 2> $O/r2d_bench_${N}gpu.err; echo "bench rc=$?"
grep -v "NCCL INFO" $O/r2d_bench_${N}gpu.err | tail -25
grep -c "NCCL INFO" $O/r2d_bench_${N}gpu.err; grep "nranks" $O/r2d_bench_${N}gpu.err | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
  tools/multi_gpu_search_check.py > $O/r2d_search_check_${N}gpu.log 2>&1; echo "search check rc=$?"; grep -v "NCCL INFO" $O/r2d_search_check_${N}gpu.log | tail -8
