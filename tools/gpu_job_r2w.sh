#!/bin/bash
# round-2 GPU job W: full GPU suite (per-test timeout), default bench line, ncu --set full of the seeded kernels as shipped
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > $O/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2w_pytest.log; tail -3 $O/r2w_pytest.log
timeout 900 python bench.py > $O/r2w_bench_1gpu.json 2> $O/r2w_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2w_bench_1gpu.json').read().strip().splitlines()[-1])
print('value',l['value'],'e2e',l['e2e']['value'],'ratio',l['e2e']['value']/l['value'],'roof',l['roofline']['frac'])
s=l['sub_records']['config4_1M_docs_seeded_one_gpu']; print('config4 one gpu', s['value'], s['roofline']['frac'])
print('pbs',l['pbs']['value'],[ (r['batch'],round(r['pbs_ms'],3)) for r in l['pbs']['by_batch']])
print('cpu',l['cpu_baseline']['value'],l['cpu_baseline']['cores'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lwe_encrypt_seeded|lincomb_seeded" \
    -s 2 -c 2 -o $O/r2w_e2e_seeded python tools/e2e_ncu.py 1000 3 > $O/r2w_ncu.log 2>&1; tail -1 $O/r2w_ncu.log
