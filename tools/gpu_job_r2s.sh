#!/bin/bash
# round-2 GPU job S: dispatcher with the pair kernel (PBS tests + sweep), ncu --set full of pbs_kernel_mb2_pair
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_pbs.py tests/test_gpu_compare.py -m gpu -x -q > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s_pytest.log; tail -3 $O/r2s_pytest.log
timeout 600 python tools/pbs_batch_sweep.py 1 2 16 37 74 75 148 296 444 592 612 666 740 1184 4736 > $O/r2s_pbs_sweep.txt 2>&1; cat $O/r2s_pbs_sweep.txt
PBS_PAIR=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pbs_kernel_mb2_pair -c 1 -o $O/r2s_pbs_pair \
    python tools/pbs_profile.py 74 2 > $O/r2s_ncu.log 2>&1
tail -2 $O/r2s_ncu.log
