cp fourierdiffusion_b200/libfdiff_b200.so /tmp/lib_keep.so
for round in 1 2; do
for v in "$@"; do
  cp tools/variants/$v.so fourierdiffusion_b200/libfdiff_b200.so
  timeout 300 python bench.py --steps 2 --warmup 3 --diffusion-steps 100 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('== $v', round(d['value']*0.1,2), r['families_us_per_launch'])"
done
done
cp /tmp/lib_keep.so fourierdiffusion_b200/libfdiff_b200.so
