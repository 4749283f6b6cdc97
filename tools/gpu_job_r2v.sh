#!/bin/bash
# round-2 GPU job V: seeded dot product variants (per-test timeouts: a hang costs minutes)
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_search.py tests/test_gpu_reference_replay.py -m gpu -x -q --timeout 90 > $O/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2v_pytest.log; tail -3 $O/r2v_pytest.log
for lib in fhe_icp_b200/libfhe_b200.so build_ab/*.so; do [ -f "$lib" ] || continue; echo "== $lib"; FHE_B200_LIB=$lib timeout 300 python tools/e2e_profile.py 1000 2>&1 | grep -E "run seeded|e2e seeded"; done > $O/r2v_e2e.txt 2>&1; cat $O/r2v_e2e.txt
