mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_capi.py tests/test_dropin_cpp.py -m gpu -q -x -p no:cacheprovider -k "sharded" > gpurun_out/r2_sharded_last.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r2_sharded_last.log
