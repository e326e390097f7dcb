# quick GPU loop: error report, parity tests, short bench
python tools/err_report.py > gpurun_out/err_report.txt 2>&1; tail -20 gpurun_out/err_report.txt
python tools/debug_ffn.py > gpurun_out/debug_ffn.txt 2>&1; tail -6 gpurun_out/debug_ffn.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
r=d['roofline']
print('value',d['value'],'e2e',d['e2e']['value'],'ffn ms',r['avg_ms_per_launch'],'frac',r['frac'],'attn ms',r['attention']['avg_ms_per_launch'], r['families_ms'])
PY
