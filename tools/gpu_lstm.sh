# LSTM (cfg 4) loop: parity tests, then the cfg4 bench line
timeout 600 python -m pytest tests -m gpu -x -q -k "lstm or golden or traj or score" > gpurun_out/pytest_lstm.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_lstm.log
timeout 600 python bench.py --config cfg4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 rc=$?"; tail -2 gpurun_out/bench_cfg4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg4.json').read().strip().splitlines()[-1])
print(d['config']['workload'], '| value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), 'launches', d['gpu_launches'])
print({k:v for k,v in (d.get('roofline') or {}).items() if k in ('kernel','achieved','frac','avg_ms_per_diffusion_step','cycles_per_recurrence_step')})
PY
