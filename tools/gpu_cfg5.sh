python bench.py --config cfg5 --steps 1 --warmup 3 --diffusion-steps 100 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 rc=$?"; tail -3 gpurun_out/bench_cfg5.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg5.json').read().strip().splitlines()[-1])
print(d['config']['workload'], '| value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'ms/step', round(d['ms_per_step'],1), (d.get('roofline') or {}).get('families_ms'))
PY
