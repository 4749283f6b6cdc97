#!/bin/bash
# round-2 GPU job G: evidence for the headline -- bench line, ncu launch list of the same command, ncu --set full of the
# lincomb kernel and of the new seeded kernels; full GPU test-suite
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2g_pytest.log; tail -3 $O/r2g_pytest.log
python bench.py --steps 20 --warmup 5 > $O/r2g_bench_1gpu.json 2> $O/r2g_bench_1gpu.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-extras --no-sub-records --no-cpu-baseline > $O/r2g_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2g_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-sub-records --no-cpu-baseline > $O/r2g_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lincomb_kernel|similarity_decrypt" -s 6 -c 2 -o $O/r2g_lincomb \
    python bench.py --steps 2 --warmup 3 --no-extras --no-sub-records --no-cpu-baseline > $O/r2g_ncu2.log 2>&1
python tools/e2e_ncu.py 1000 3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"quantize_body_noise|lwe_encrypt_seeded|lincomb_seeded|similarity_decrypt" \
    -s 4 -c 4 -o $O/r2g_e2e_seeded python tools/e2e_ncu.py 1000 3 > $O/r2g_ncu3.log 2>&1
ls -la $O | grep r2g
