#!/bin/bash
# round-2 GPU job 3D: pbs_kernel_mb2<1,4> with a one-time start offset for the second warp of every scheduler (A/B builds)
O=gpurun_out
for lib in fhe_icp_b200/libfhe_b200.so build_ab/libfhe_ms_*.so; do [ -f "$lib" ] || continue; echo "== $lib"; FHE_B200_LIB=$lib SWEEP_ONLY=dispatch timeout 120 python tools/pbs_batch_sweep.py 592 1184 2>&1 | tail -2; done > $O/r3d_mb2_skew.txt 2>&1; cat $O/r3d_mb2_skew.txt
