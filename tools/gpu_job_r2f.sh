#!/bin/bash
# round-2 GPU job F: ring depth of the PBS kernels (small batches through the split kernel, large through mb2)
O=gpurun_out
python -m pytest tests/test_gpu_pbs.py tests/test_gpu_search.py -m gpu -x -q > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2f_pytest.log; tail -3 $O/r2f_pytest.log
for v in default ps4 ps8 mb5; do
  L=fhe_icp_b200/libfhe_b200.so; [ $v = default ] || L=build_ab/libfhe_$v.so
  echo "== $v"
  for B in 1 16 148 296 592 1184 4736; do
    FHE_B200_LIB=$L PBS_MB2=1 python tools/pbs_profile.py $B 3 | tail -2 | tr '\n' ' '; echo
  done
done > $O/r2f_pbs_ring.txt 2>&1
cat $O/r2f_pbs_ring.txt
