#!/bin/bash
# round-2 GPU job K: first run of the four-warps-per-polynomial latency kernel (pbs_wide.cu)
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_pbs.py -m gpu -x -q -k "wide" > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2k_pytest.log; tail -15 $O/r2k_pytest.log
timeout 600 python tools/pbs_batch_sweep.py 1 16 74 148 296 592 1184 > $O/r2k_pbs_sweep.txt 2>&1; cat $O/r2k_pbs_sweep.txt
