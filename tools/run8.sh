TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
$TR 29601 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu_v2.json 2> gpurun_out/bench_8gpu_v2.err
$TR 29602 bench.py --gpus 8 --workload train --batch 2 --steps 10 --warmup 3 > gpurun_out/bench_train_8gpu_v2.json 2> gpurun_out/bench_train_8gpu_v2.err
FCWDM_DDP_OVERLAP=0 $TR 29603 bench.py --gpus 8 --workload train --batch 2 --steps 10 --warmup 3 > gpurun_out/bench_train_8gpu_nooverlap.json 2>/dev/null
NCCL_MAX_NCHANNELS=4 $TR 29604 bench.py --gpus 8 --workload train --batch 2 --steps 10 --warmup 3 > gpurun_out/bench_train_8gpu_4ch.json 2>/dev/null
for f in bench_8gpu_v2 bench_train_8gpu_v2 bench_train_8gpu_nooverlap bench_train_8gpu_4ch; do echo $f; tail -n 1 gpurun_out/$f.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"; done
