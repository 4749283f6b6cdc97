#!/bin/bash
# round-2 GPU job U (N GPUs): bench.py exactly as the driver launches it at N > 1, both arms, plus the multi-GPU search check
O=gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 20 --warmup 5 > $O/r2u_bench_${N}gpu.json 2> $O/r2u_bench_${N}gpu.err; echo "bench rc=$?"
grep -v "NCCL INFO" $O/r2u_bench_${N}gpu.err | tail -8
python - <<PY
import json
l=json.loads(open('$O/r2u_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('N',l['n_gpus'],'value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'scaling',l['scaling'],'workload',l['config'].get('workload'))
for k,v in l.get('sub_records',{}).items(): print('  sub',k,v.get('value'),v.get('ms_per_step'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
  tools/multi_gpu_search_check.py > $O/r2u_search_check_${N}gpu.log 2>&1; echo "search check rc=$?"; grep -v "NCCL INFO" $O/r2u_search_check_${N}gpu.log | tail -4
