# A/B of the "stack_lanes" option inside ONE gpurun call (boxes differ by ~3 %): whole-sampler series/s at cfg 2 and cfg 3 with 1 / 2 / 3 lanes.
# usage: bash tools/ab_stack_lanes.sh     -> gpurun_out/ab_stack_lanes.txt
out=gpurun_out/ab_stack_lanes.txt; : > $out
for round in 1 2; do
for lanes in 1 2 3; do
  for cfg in cfg2 cfg3; do
    FD_STACK_LANES=$lanes timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --profile-stride 0 2>/dev/null \
      | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('round $round lanes $lanes $cfg', d['value'], d['e2e']['value'], d['ms_per_step'], d.get('gpu_launches'))" >> $out
  done
done
done
cat $out
