#!/bin/bash
# 2-GPU box: the NCCL tests the 1-GPU driver box skips, then the default bench line at N = 2 as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_ddp_gpu.py tests/test_trainloop_gpu.py -x -q -rs > gpurun_out/r02_gputest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_gputest_2gpu.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench rc=$?"
for w in fp32 bf16; do
FCWDM_DDP_GRAD_DTYPE=$w timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_train_2gpu_$w.json 2> gpurun_out/r02_bench_train_2gpu_$w.err; echo "bench train $w rc=$?"
done
python - <<'PY'
import json
for n in ("bench_2gpu","bench_train_2gpu_fp32","bench_train_2gpu_bf16"):
    try:
        d=json.load(open(f"gpurun_out/r02_{n}.json"))
        t=(d.get("secondary") or {}).get("train") or {}
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"], "| train:", t.get("value"), t.get("clocks"))
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r02_bench_2gpu.err
