#!/bin/bash
# round-2 GPU job T: ncu --set full of glwe_dot_kernel (the shipped, post-prefetch build); ncu launch list of the bench command
O=gpurun_out
timeout 300 python tools/packed_profile.py 262144 2 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:glwe_dot_kernel -s 1 -c 1 -o $O/r2t_glwe_dot \
    python tools/packed_profile.py 262144 2 > $O/r2t_ncu.log 2>&1; tail -2 $O/r2t_ncu.log
python bench.py --steps 2 --warmup 3 --no-extras --no-sub-records --no-cpu-baseline > $O/r2t_b.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2t_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-sub-records --no-cpu-baseline > $O/r2t_ncu1.log 2>&1; tail -1 $O/r2t_ncu1.log | cut -c1-300
