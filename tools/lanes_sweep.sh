for l in 2 3 4; do
FD_LANES=$l python bench.py --no-cpu-baseline --steps 2 --warmup 3 --profile-stride 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lanes $l value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
done
