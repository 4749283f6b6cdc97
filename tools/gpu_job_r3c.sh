#!/bin/bash
# round-2 GPU job 3C (1 GPU): what the driver runs at round end -- full GPU suite, smoke(), the default bench line
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > $O/r3c_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r3c_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r3c_bench_1gpu.json 2> $O/r3c_bench_1gpu.err; echo "bench rc=$?"; tail -3 $O/r3c_bench_1gpu.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r3c_bench_1gpu.json').read().strip().splitlines()[-1])
print('value',l['value'],'e2e',l['e2e']['value'],[round(b/1e6,2) for b in l['e2e']['blocks']],'roof',l['roofline']['frac'])
print('pbs',l['pbs']['value'],[(r['batch'],round(r['pbs_ms'],3)) for r in l['pbs']['by_batch']])
print('pbs cpu',l['pbs'].get('cpu_baseline'))
print('sharded',[(w['scaling'],round(w['ks_pbs_per_sec'])) for w in l['pbs_sharded']['workloads']], l['pbs_sharded']['key_broadcast'])
print('cpu',l['cpu_baseline']['value'],l['cpu_baseline']['cores'])
PY
