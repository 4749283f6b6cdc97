#!/bin/bash
# round-2 GPU job X: A/B builds of the seeded dot product (registers per thread / unroll / round keys from the constant bank)
O=gpurun_out
for lib in fhe_icp_b200/libfhe_b200.so build_ab/*.so; do [ -f "$lib" ] || continue; echo "== $lib"; FHE_B200_LIB=$lib timeout 200 python tools/e2e_profile.py 1000 2>&1 | grep -E "encrypt seeded|run seeded|e2e seeded"; done > $O/r2x_seeded_ab.txt 2>&1; cat $O/r2x_seeded_ab.txt
