#!/bin/bash
# round-2 GPU job B: parity of the reworked seeded kernels + A/B of their build variants + launch list
O=gpurun_out
python -m pytest tests/test_gpu_linear.py tests/test_gpu_search.py tests/test_gpu_pbs.py -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
tail -3 $O/r2b_pytest.log
for v in default u2 u3 c4 c6u2; do
  echo "== $v"
  if [ $v = default ]; then python tools/e2e_profile.py 1000; else FHE_B200_LIB=build_ab/libfhe_$v.so python tools/e2e_profile.py 1000; fi
done > $O/r2b_ab.txt 2>&1
cat $O/r2b_ab.txt
for v in default u2 u3 c4 c6u2; do
  L=fhe_icp_b200/libfhe_b200.so; [ $v = default ] || L=build_ab/libfhe_$v.so
  FHE_B200_LIB=$L ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2b_launches_$v.csv python tools/e2e_ncu.py 1000 3 > /dev/null 2>&1
  echo "== $v"; grep -E "lwe_encrypt_seeded|lincomb_seeded" $O/r2b_launches_$v.csv | awk -F'","' '{print $5, $NF}' | tail -4
done
