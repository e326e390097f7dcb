for st in 0 3000 6000 9000; do
for lanes in 1 2; do
FD_LANES=$lanes FD_FFN_STAGGER_NS=$st python bench.py --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('stagger $st lanes $lanes value', round(d['value'],1), 'ffn us', round(r['avg_ms_per_launch']*1e3,1), 'attn us', round(r['attention']['avg_ms_per_launch']*1e3,1))"
done; done
