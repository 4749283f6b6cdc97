#!/bin/bash
# round-2 GPU job Y: phase skew between the ciphertexts of a CTA of pbs_kernel_mb2<1,4> (A/B builds)
O=gpurun_out
for lib in fhe_icp_b200/libfhe_b200.so build_ab/*.so; do [ -f "$lib" ] || continue; echo "== $lib"; FHE_B200_LIB=$lib SWEEP_ONLY=dispatch timeout 120 python tools/pbs_batch_sweep.py 592 1184 2>&1 | tail -2; done > $O/r2y_mb2_skew.txt 2>&1; cat $O/r2y_mb2_skew.txt
