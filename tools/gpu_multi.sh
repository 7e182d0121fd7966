#!/bin/bash
# multi-GPU bench under torchrun with the library's exchange: gpurun --gpus N -- bash tools/gpu_multi.sh N [tag]
set -u
N=${1:-2}; TAG=${2:-r02f}
mkdir -p gpurun_out
env BFC_DIST_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/${TAG}_n$N.err | tail -8
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_n$N.json").read().strip().splitlines()[-1])
print(d["metric"], "N", d["n_gpus"], round(d["value"],2), "ms/step", round(d["ms_per_step"],1), "e2e", d["e2e"] and round(d["e2e"]["value"],2), d["e2e"] and d["e2e"].get("equals_resident_result"))
print(" exchange", d.get("exchange"))
for k,v in sorted(d["roofline"]["kernels"].items(), key=lambda kv:-kv[1]["ms"]):
    print("   %-14s %8.1f ms/step share %.3f" % (k, v["ms"]/d["steps"], v["share_of_step"]))
PY
ls gpurun_out/bench_error_rank*.txt 2>/dev/null && cat gpurun_out/bench_error_rank*.txt | head -30
