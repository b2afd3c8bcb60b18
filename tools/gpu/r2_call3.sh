mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v7.json 2> gpurun_out/r2_bench_v7.err; echo "bench rc $?"
tail -5 gpurun_out/r2_bench_v7.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref_v7.json 2> gpurun_out/r2_bench_ref_v7.err; echo "ref rc $?"
