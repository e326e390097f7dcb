for l0 in 0 148 160 108 96 192; do
FD_LANE0=$l0 python bench.py --no-cpu-baseline --steps 2 --warmup 3 --profile-stride 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lane0 $l0 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
done
