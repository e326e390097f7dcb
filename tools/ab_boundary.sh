# A/B of the step-boundary kernel variants inside ONE gpurun call: FD_FUSE_BOUNDARY=2 (weights in shared memory) vs 1 (constant operands)
# usage: bash tools/ab_boundary.sh [cfg ...]   (default cfg2)
out=gpurun_out/ab_boundary.txt; : > $out
for cfg in ${@:-cfg2}; do
for round in 1 2; do
for fb in 2 1; do
  FD_FUSE_BOUNDARY=$fb timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --profile-stride 50 2>/dev/null \
    | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg round $round fuse_boundary $fb', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['families_us_per_launch'])" >> $out
done
done
done
cat $out
