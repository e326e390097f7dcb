# A/B of the step-boundary kernel variants inside ONE gpurun call: FD_FUSE_BOUNDARY=2 (weights in shared memory) vs 1 (constant operands),
# (FD_SB_TOK64 was a 64-token-per-CTA variant of the latter: measured the same and removed)
out=gpurun_out/ab_boundary.txt; : > $out
for round in 1 2; do
for v in "2 0" "1 0"; do
  set -- $v
  FD_FUSE_BOUNDARY=$1 FD_SB_TOK64=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --profile-stride 50 2>/dev/null \
    | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('round $round fuse_boundary $1 tok64 $2', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['families_us_per_launch'])" >> $out
done
done
cat $out
