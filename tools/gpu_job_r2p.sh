#!/bin/bash
# round-2 GPU job P: the dispatcher with the wide latency kernel (full GPU PBS / compare / search tests), batch sweep,
# ncu --set full of the shipped wide kernel
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_pbs.py tests/test_gpu_compare.py tests/test_gpu_packed.py -m gpu -x -q > $O/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2p_pytest.log; tail -4 $O/r2p_pytest.log
timeout 600 python tools/pbs_batch_sweep.py 1 16 74 148 296 444 592 612 740 1036 1184 4736 > $O/r2p_pbs_sweep.txt 2>&1; cat $O/r2p_pbs_sweep.txt
PBS_WIDE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pbs_kernel_mb2_wide -c 1 -o $O/r2p_pbs_wide \
    python tools/pbs_profile.py 148 2 > $O/r2p_ncu.log 2>&1
tail -2 $O/r2p_ncu.log
