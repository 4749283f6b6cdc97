#!/bin/bash
# round-2 GPU job Q: full GPU suite + default bench line + smoke after the e2e / PBS dispatcher changes
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2q_pytest.log; tail -3 $O/r2q_pytest.log
timeout 900 python bench.py > $O/r2q_bench_1gpu.json 2> $O/r2q_bench_1gpu.err; echo "bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2q_bench_1gpu.json').read().strip().splitlines()[-1])
print('value',l['value'],'e2e',l['e2e']['value'],'ratio',l['e2e']['value']/l['value'],'roof',l['roofline']['frac'])
print('pbs',l['pbs']['value'],[ (r['batch'],round(r['pbs_ms'],3)) for r in l['pbs']['by_batch']])
PY
