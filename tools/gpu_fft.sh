# dft / idft loop: parity tests, then the HBM-fraction table
timeout 600 python -m pytest tests -m gpu -x -q -k "dft or fourier or time_domain" > gpurun_out/pytest_fft.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_fft.log
timeout 300 python tools/fft_bench.py > gpurun_out/fft_bench.txt 2>&1; cat gpurun_out/fft_bench.txt
