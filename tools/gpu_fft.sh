# dft / idft loop: parity tests, then the HBM-fraction table
timeout 600 python -m pytest tests -m gpu -x -q -k "dft or fourier or time_domain" > gpurun_out/pytest_fft.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_fft.log
timeout 300 python tools/fft_bench.py > gpurun_out/fft_bench.txt 2>&1; cat gpurun_out/fft_bench.txt
if [ -n "$FFT_EXTRA" ]; then FD_FFT_CW4=1 timeout 300 python tools/fft_bench.py 2>&1 | grep cfg5; FD_FFT_CW4=1 timeout 600 python -m pytest tests -m gpu -x -q -k "dft" 2>&1 | tail -2; fi
