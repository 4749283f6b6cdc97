mkdir -p gpurun_out
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534"
for i in 1 2; do timeout 300 $TR2 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/b2_push.json 2> gpurun_out/b2_push.err; echo "b2 push rc=$?"; cut -c1-200 gpurun_out/b2_push.json; done
timeout 300 $TR2 tools/multi_gpu_search_check.py > gpurun_out/mg2_check.log 2>&1; echo "check2 rc=$?"; tail -1 gpurun_out/mg2_check.log
