mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_peer_exchange.py tests/test_sharded_capi.py tests/test_dropin_cpp.py -m gpu -q -x -p no:cacheprovider -k "peer or sharded" > gpurun_out/r2_n8_tests.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r2_n8_tests.log
timeout 900 python tools/sharded_e2e.py c4 c4_400k > gpurun_out/r2_sharded_e2e.jsonl 2> gpurun_out/r2_sharded_e2e.err; echo "e2e rc $?"; cut -c1-420 gpurun_out/r2_sharded_e2e.jsonl; tail -3 gpurun_out/r2_sharded_e2e.err
