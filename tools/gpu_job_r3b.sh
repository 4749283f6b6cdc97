#!/bin/bash
# round-2 GPU job 3B (N GPUs): bench.py exactly as the driver launches it at N > 1 (the new pbs_sharded record runs on every rank)
O=gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 20 --warmup 5 > $O/r3b_bench_${N}gpu.json 2> $O/r3b_bench_${N}gpu.err; echo "bench rc=$?"
grep -v "NCCL INFO" $O/r3b_bench_${N}gpu.err | tail -12
python - <<PY
import json
l=json.loads(open('$O/r3b_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('N',l['n_gpus'],'value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'scaling',l['scaling'])
for k,v in l.get('sub_records',{}).items(): print('  sub',k,v.get('value'),v.get('ms_per_step'))
print('pbs',l['pbs']['value'])
ps=l['pbs_sharded']
for w in ps['workloads']: print('  sharded',w['scaling'],w['batch_total'],w['per_rank'],'ks+pbs/s',round(w['ks_pbs_per_sec']),'round trip/s',round(w['round_trip_per_sec']),'frac',w['frac_of_fp64_term'],w['all_correct'])
print('  key broadcast',ps['key_broadcast'])
PY
