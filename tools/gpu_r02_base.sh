# round-2 evidence run: GPU tests, smoke, full bench, reference arm, ncu launch list + one --set full capture of the stack kernel
TAG=${1:-r02}
set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --diffusion-steps 12 --no-cpu-baseline --no-other-configs --profile-stride 0 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'encoder_stack' -s 2 -c 1 -f -o gpurun_out/${TAG}_full \
    python tools/profile_layer.py > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"; tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/${TAG}_*
