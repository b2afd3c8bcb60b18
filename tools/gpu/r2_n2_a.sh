mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_peer_exchange.py tests/test_sharded.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_n2_tests.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/r2_n2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_strong.json 2> gpurun_out/r2_bench_n2_strong.err; echo "bench rc $?"; tail -5 gpurun_out/r2_bench_n2_strong.err
