mkdir -p gpurun_out
timeout 300 python tools/window_variants.py c4 > gpurun_out/plain3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:window_kernel -s 10 -c 2 -o gpurun_out/r2_prof_c4_v7 python tools/window_variants.py c4 > gpurun_out/ncu_f7.log 2>&1; echo "ncu full rc $?"
