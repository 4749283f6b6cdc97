#!/bin/bash
# round-2 GPU job C: the reworked bench.py on one GPU (both workloads), the full GPU test-suite
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c_pytest.log; tail -3 $O/r2c_pytest.log
python bench.py --steps 20 --warmup 5 > $O/r2c_bench_1gpu.json 2> $O/r2c_bench_1gpu.err; echo "bench rc=$?"; tail -5 $O/r2c_bench_1gpu.err
python bench.py --workload config4 --total-docs 200000 --steps 5 --warmup 3 --no-extras --no-cpu-baseline --docs5-per-gpu 2000 > $O/r2c_bench_cfg4_1gpu.json 2> $O/r2c_bench_cfg4_1gpu.err; echo "bench cfg4 rc=$?"; tail -5 $O/r2c_bench_cfg4_1gpu.err
