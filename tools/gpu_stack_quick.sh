# quick loop for the encoder-stack kernel: parity tests, per-phase cycle breakdown, short bench, streaming-attention timing
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
LAGS=-1 PHASES=1 timeout 300 python tools/stack_probe.py 2>&1 | tail -4 | cut -c1-900
timeout 600 python bench.py --steps 2 --warmup 3 --diffusion-steps 200 --no-cpu-baseline --no-other-configs > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
r=d['roofline']
print('value x0.2', d['value']*0.2, 'stack us', r['families_us_per_launch'], 'tasks', r.get('tasks'))
PY
timeout 300 python tools/time_lengths.py 2>&1 | tail -6
