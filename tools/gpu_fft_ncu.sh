# ncu --set full of the dft / idft kernels at the cfg2 (x64 batches) and cfg5 shapes
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rfft' -c 6 -f -o gpurun_out/${1:-r02}_fft_full python tools/fft_profile.py > gpurun_out/fft_ncu.log 2>&1; echo rc=$?; tail -3 gpurun_out/fft_ncu.log
