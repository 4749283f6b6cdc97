mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_search.py -m gpu -x -q ) > gpurun_out/t_lin.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_lin.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err; echo "b1 rc=$?"; cut -c1-200 gpurun_out/b1.json
