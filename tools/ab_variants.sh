# A/B of library variants inside ONE gpurun call (boxes of the pool differ by ~3 %): tools/variants/<name>.so are copied over the library in turn
# usage: bash tools/ab_variants.sh v0 v1 v2 ...   (STREAM=1: also the per-token time at L = 4096 through the streaming attention kernel)
cp fourierdiffusion_b200/libfdiff_b200.so /tmp/lib_keep.so
for round in 1 2; do
for v in "$@"; do
  cp tools/variants/$v.so fourierdiffusion_b200/libfdiff_b200.so
  echo "== $v (round $round)"; LAGS=-1 timeout 200 python tools/stack_probe.py 2>&1 | tail -1 | cut -c1-40
  if [ -n "$STREAM" ]; then timeout 200 python tools/time_lengths.py 2>&1 | tail -1; fi
done
done
cp /tmp/lib_keep.so fourierdiffusion_b200/libfdiff_b200.so
