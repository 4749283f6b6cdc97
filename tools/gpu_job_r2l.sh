#!/bin/bash
# round-2 GPU job L: ncu --set full of the wide latency kernel (batch 148 = one CTA per SM)
O=gpurun_out
PBS_WIDE=1 timeout 300 python tools/pbs_profile.py 148 2 && \
PBS_WIDE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pbs_kernel_mb2_wide -c 1 -o $O/r2l_pbs_wide \
    python tools/pbs_profile.py 148 2 > $O/r2l_ncu.log 2>&1
tail -3 $O/r2l_ncu.log
