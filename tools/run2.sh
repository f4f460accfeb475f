TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$TR 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_v3.json 2>/dev/null
$TR 29612 bench.py --gpus 2 --workload train --batch 2 --steps 10 --warmup 3 > gpurun_out/bench_train_2gpu_v2.json 2>/dev/null
for f in bench_2gpu_v3 bench_train_2gpu_v2; do tail -n 1 gpurun_out/$f.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['metric'], d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"; done
