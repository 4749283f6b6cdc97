#!/bin/bash
# round-2 GPU job E (8 GPUs): bench.py as the driver launches it at N=8, plus shard-weight variants of the headline
O=gpurun_out
N=${1:-8}
run() { tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 "$@" > $O/r2e_bench_${N}gpu_$tag.json 2> $O/r2e_bench_${N}gpu_$tag.err; echo "bench $tag rc=$?"
  grep -v "NCCL INFO" $O/r2e_bench_${N}gpu_$tag.err | grep -v "^\*\*\*\|OMP_NUM_THREADS" | tail -15
  python - <<PY
import json
l=json.loads(open("$O/r2e_bench_${N}gpu_$tag.json").read().strip().splitlines()[-1])
print("$tag", "value", round(l["value"]), "ms", round(l["ms_per_step"],3), "e2e", round(l["e2e"]["value"]), "res", round(l["e2e"].get("resident_collection_value",0)), "rho", l["client_rho"], "shards", l["shards"][:3])
for k,v in l.get("sub_records",{}).items(): print("  ", k, round(v["value"]), v["ms_per_step"], v.get("shards",[None])[:2], v.get("roofline",{}).get("frac"))
PY
}
run full
run equal --client-rho 0 --no-sub-records --no-extras
run pad3 --client-rho-pad 3 --no-sub-records --no-extras
run weak_equal --workload config2 --client-rho 0 --no-sub-records --no-extras
run weak_pad3 --workload config2 --client-rho-pad 3 --no-sub-records --no-extras
grep -c "NCCL INFO" $O/r2e_bench_${N}gpu_full.err; grep "nranks" $O/r2e_bench_${N}gpu_full.err | head -3
