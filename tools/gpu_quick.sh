# quick GPU loop: error report, parity tests, short bench, attention timeline
python tools/err_report.py > gpurun_out/err_report.txt 2>&1; grep tensor-core gpurun_out/err_report.txt || tail -5 gpurun_out/err_report.txt
FD_ATTN_BOUNDED=0 python tools/err_report.py 2>&1 | grep tensor-core | sed "s/^/exact-softmax /"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
r=d['roofline']
print('value',d['value'],'e2e',d['e2e']['value'],'ffn ms',r['avg_ms_per_launch'],'frac',r['frac'],'attn ms',r['attention']['avg_ms_per_launch'], r['families_ms'])
PY
python tools/attn_timeline.py 256 > gpurun_out/tl256.txt 2>&1
