set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/r2_gputest_v5.log 2>&1; echo "pytest rc $?"
tail -25 gpurun_out/r2_gputest_v5.log
timeout 600 python tools/window_variants.py > gpurun_out/r2_variants_v5.jsonl 2> gpurun_out/r2_variants_v5.err; echo "variants rc $?"
cat gpurun_out/r2_variants_v5.jsonl; tail -5 gpurun_out/r2_variants_v5.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v5.json 2> gpurun_out/r2_bench_v5.err; echo "bench rc $?"
tail -3 gpurun_out/r2_bench_v5.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_c4_v5.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc $?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:window_kernel -s 6 -c 2 -o gpurun_out/r2_prof_c4_v5 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1; echo "ncu full rc $?"
