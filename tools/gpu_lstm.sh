python tools/err_report.py 2>&1 | grep lstm
timeout 600 python -m pytest tests -m gpu -x -q -k "lstm or mimic" 2>&1 | tail -3
for tc in 1 0; do FD_LSTM_TC=$tc python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tc $tc cfg4 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],1))"; done
