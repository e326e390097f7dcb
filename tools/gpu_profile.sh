# ncu evidence for profiles/: launch list of a short bench run + one --set full capture of the two layer kernels
TAG=${1:-r01}
export FD_LANES=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --diffusion-steps 12 --no-cpu-baseline --profile-stride 0 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ffn_ln|attention_fused' -s 22 -c 2 -f -o gpurun_out/${TAG}_full \
    python tools/profile_layer.py > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/${TAG}_*
