#!/bin/bash
# 8-GPU box: the default bench line as the driver launches it at N = 8 (sampling, batch 8 and the DDP training block), the
# reference arm, and the training workload with the bf16 wire for comparison
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu_v2.json 2> gpurun_out/r02_bench_8gpu_v2.err; echo "bench rc=$?"
for w in fp32 bf16; do
FCWDM_DDP_GRAD_DTYPE=$w timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload train --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_train_8gpu_v2_$w.json 2> gpurun_out/r02_bench_train_8gpu_v2_$w.err; echo "bench train $w rc=$?"
done
python - <<'PY'
import json
for n in ("bench_8gpu_v2","bench_train_8gpu_v2_fp32","bench_train_8gpu_v2_bf16"):
    try:
        for line in open(f"gpurun_out/r02_{n}.json"):
            if line.startswith("{"):
                d=json.loads(line)
                t=(d.get("secondary") or {}).get("train") or {}
                b=(d.get("secondary") or {}).get("batch8") or {}
                print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"], "| train:", t.get("value"), (t.get("e2e") or {}).get("value"), t.get("clocks"), "| b8:", b.get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r02_bench_8gpu_v2.err
