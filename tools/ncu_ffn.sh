timeout 600 ncu --set full --clock-control none -k regex:'ffn_ln' -s 6 -c 1 -f -o gpurun_out/ffn128 python tools/profile_layer.py > gpurun_out/ffn128_ncu.log 2>&1; echo rc=$?
FD_FFN_TILE=256 timeout 600 ncu --set full --clock-control none -k regex:'ffn_ln' -s 6 -c 1 -f -o gpurun_out/ffn256 python tools/profile_layer.py > gpurun_out/ffn256_ncu.log 2>&1; echo rc=$?
