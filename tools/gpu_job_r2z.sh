#!/bin/bash
# round-2 GPU job Z: initial phase offsets between the ciphertext slots of glwe_dot_kernel (A/B builds)
O=gpurun_out
for lib in fhe_icp_b200/libfhe_b200.so build_ab/*.so; do [ -f "$lib" ] || continue; echo "== $lib"; FHE_B200_LIB=$lib timeout 120 python tools/packed_profile.py 262144 3 2>&1 | tail -3; done > $O/r2z_glwe_skew.txt 2>&1; cat $O/r2z_glwe_skew.txt
