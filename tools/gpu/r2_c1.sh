mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_capi.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_sharded_capi.log 2>&1; echo "pytest rc $?"; tail -8 gpurun_out/r2_sharded_capi.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/plain_c1.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_c4_c1.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc $?"
bash tools/gpu/r2_prof.sh c1 c4
