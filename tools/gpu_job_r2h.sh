#!/bin/bash
# round-2 GPU job H: chunk-pipelined predict_host_seeded
O=gpurun_out
python -m pytest tests/test_gpu_linear.py tests/test_gpu_search.py tests/test_gpu_reference_replay.py -m gpu -x -q > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2h_pytest.log; tail -3 $O/r2h_pytest.log
python tools/e2e_profile.py 1000 > $O/r2h_e2e.txt 2>&1; cat $O/r2h_e2e.txt
python tools/e2e_profile.py 300 | tail -2; python tools/e2e_profile.py 5000 | tail -2
python bench.py --steps 20 --warmup 5 --no-extras --no-sub-records --no-cpu-baseline | python -c "
import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',l['value'],'e2e',l['e2e']['value'], 'ratio', l['e2e']['value']/l['value'])"
