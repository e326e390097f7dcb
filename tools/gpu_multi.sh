# multi-GPU check: bench.py under torchrun exactly as the driver launches it
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
tail -3 gpurun_out/bench_n$N.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'e2e',d['e2e'],'ms/step',d['ms_per_step'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 1 --warmup 1 | tail -1 | cut -c1-300
