#!/bin/bash
O=gpurun_out
for v in nct2s1 nct2r; do
  L=build_ab/libfhe_$v.so
  echo "== $v"
  for B in 1184; do
    FHE_B200_PBS_DEBUG=1 FHE_B200_LIB=$L PBS_MB2=1 python tools/pbs_profile.py $B 2 2>&1 | tail -3 | tr '\n' ' '; echo
  done
done > $O/r2i_pbs_nct2b.txt 2>&1
cat $O/r2i_pbs_nct2b.txt
