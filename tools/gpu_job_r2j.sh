#!/bin/bash
# round-2 GPU job J: e2e fixed overhead (pinned inputs, zero-copy result) + regression of the touched paths
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_pbs.py tests/test_gpu_search.py -m gpu -x -q > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2j_pytest.log; tail -3 $O/r2j_pytest.log
for m in 0 4 2 1; do FHE_B200_E2E_MODE=$m timeout 300 python tools/e2e_ab.py 1000; done > $O/r2j_e2e_ab.txt 2>&1
cat $O/r2j_e2e_ab.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-sub-records --no-cpu-baseline | python -c "
import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',l['value'],'e2e',l['e2e']['value'], 'ratio', l['e2e']['value']/l['value'])"
