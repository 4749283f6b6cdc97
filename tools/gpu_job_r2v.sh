#!/bin/bash
# round-2 GPU job V: seeded dot product without per-word selects in the inner loop (per-test timeouts: a hang costs minutes)
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_search.py tests/test_gpu_reference_replay.py -m gpu -x -q --timeout 90 > $O/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2v_pytest.log; tail -3 $O/r2v_pytest.log
timeout 300 python tools/e2e_profile.py 1000 > $O/r2v_e2e.txt 2>&1; tail -3 $O/r2v_e2e.txt
