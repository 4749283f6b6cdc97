#!/bin/bash
# round-2 GPU job 3A (1 GPU): the ShardedBootstrap GPU test, then the default bench line with the new pbs_sharded record
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_pbs.py -m gpu -x -q --timeout 120 -k "sharded_bootstrap" > $O/r3a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r3a_pytest.log
timeout 900 python bench.py > $O/r3a_bench_1gpu.json 2> $O/r3a_bench_1gpu.err; echo "bench rc=$?"; tail -5 $O/r3a_bench_1gpu.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r3a_bench_1gpu.json').read().strip().splitlines()[-1])
print('value',l['value'],'e2e',l['e2e']['value'],'roof',l['roofline']['frac'])
print('pbs',l['pbs']['value'])
print('pbs_sharded',json.dumps(l['pbs_sharded'])[:1500])
PY
