#!/bin/bash
# round-2 GPU job R: first run of the two-CTA cluster kernel (tight timeouts: a lost mbarrier signal would hang it)
O=gpurun_out
timeout 120 python -m pytest tests/test_gpu_pbs.py -m gpu -x -q -k "latency and pair and toy" > $O/r2r_pytest.log 2>&1; echo "toy rc=$?" >> $O/r2r_pytest.log
timeout 180 python -m pytest tests/test_gpu_pbs.py -m gpu -x -q -k "latency" >> $O/r2r_pytest.log 2>&1; echo "all rc=$?" >> $O/r2r_pytest.log; tail -12 $O/r2r_pytest.log
SWEEP_ONLY=pair,wide timeout 200 python tools/pbs_batch_sweep.py 1 16 74 148 > $O/r2r_pbs_sweep.txt 2>&1; cat $O/r2r_pbs_sweep.txt
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
