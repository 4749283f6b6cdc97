#!/bin/bash
# round-2 GPU job Y: the throughput kernel pbs_kernel_mb2<1,4> (tests with per-test timeout, then timings)
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_pbs.py tests/test_gpu_compare.py -m gpu -x -q --timeout 120 2>&1 | tail -3
for lib in fhe_icp_b200/libfhe_b200.so build_ab/*.so; do [ -f "$lib" ] || continue; echo "== $lib"; FHE_B200_LIB=$lib SWEEP_ONLY=dispatch timeout 120 python tools/pbs_batch_sweep.py 592 1184 4736 2>&1 | tail -3; done > $O/r2y_mb2.txt 2>&1; cat $O/r2y_mb2.txt
