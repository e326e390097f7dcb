# A/B of the step-boundary kernel variants inside ONE gpurun call: FD_FUSE_BOUNDARY=1 (weights as constant operands) vs 2 (weights in shared memory)
out=gpurun_out/ab_boundary.txt; : > $out
for round in 1 2; do
for fb in 2 1; do
  FD_FUSE_BOUNDARY=$fb timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --profile-stride 50 2>/dev/null \
    | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('round $round fuse_boundary $fb', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['families_us_per_launch'])" >> $out
done
done
cat $out
