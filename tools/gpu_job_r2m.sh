#!/bin/bash
# round-2 GPU job M: wide kernel A/B (library variants under build_ab/ if present) + its tests (per-test timeout)
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_pbs.py -m gpu -x -q -k "latency or multibit_pbs" --timeout 60 2>&1 | tail -3
: > $O/r2m_pbs_wide_ab.txt
for lib in fhe_icp_b200/libfhe_b200.so build_ab/*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib" >> $O/r2m_pbs_wide_ab.txt
  FHE_B200_LIB=$lib SWEEP_ONLY=wide timeout 120 python tools/pbs_batch_sweep.py 1 16 148 >> $O/r2m_pbs_wide_ab.txt 2>&1
done
cat $O/r2m_pbs_wide_ab.txt
