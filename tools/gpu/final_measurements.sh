mkdir -p gpurun_out
TAG=${1:-f1}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2_gputest_$TAG.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2_gputest_$TAG.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_$TAG.json 2> gpurun_out/r2_bench_$TAG.err; echo "bench rc $?"; tail -3 gpurun_out/r2_bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref_$TAG.json 2> gpurun_out/r2_bench_ref_$TAG.err; echo "ref rc $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_$TAG.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r2_smoke_$TAG.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_c4_$TAG.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_kernel|pattern_kernel|route_kernel" -s 12 -c 4 -o gpurun_out/r2_prof_c4_$TAG -f python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_f_$TAG.log 2>&1; echo "ncu full rc $?"
