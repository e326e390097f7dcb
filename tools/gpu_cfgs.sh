for cfg in cfg3 cfg4; do
python bench.py --config $cfg --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; echo "$cfg rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$cfg.json').read().strip().splitlines()[-1])
print(d['config']['workload'], '| value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],1), (d.get('roofline') or {}).get('families_ms'))
PY
done
