#!/bin/bash
# round-2 GPU job A: tests, e2e stage times, bench line, ncu launch list + full captures of the seeded e2e kernels
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
tail -3 $O/r2a_pytest.log
python tools/e2e_profile.py 1000 > $O/r2a_e2e_stage_times.txt 2>&1; cat $O/r2a_e2e_stage_times.txt
python bench.py --steps 20 --warmup 5 > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"
python tools/e2e_ncu.py 1000 3 > $O/r2a_e2e_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2a_launches_e2e.csv \
    python tools/e2e_ncu.py 1000 3 > $O/r2a_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"quantize_body_noise|lwe_encrypt_seeded|lincomb_seeded|similarity_decrypt" \
    -s 4 -c 4 -o $O/r2a_e2e_seeded python tools/e2e_ncu.py 1000 3 > $O/r2a_ncu2.log 2>&1
ls -la $O
