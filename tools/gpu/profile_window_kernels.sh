# usage: bash tools/gpu/profile_window_kernels.sh TAG [workload]   — ncu full capture of the window kernels (count + emit) of one search
TAG=${1:-x}; WL=${2:-c4}
mkdir -p gpurun_out
timeout 300 python bench.py --workload $WL --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:window_kernel -s 8 -c 2 -o gpurun_out/r2_prof_${WL}_$TAG -f python bench.py --workload $WL --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc $?"
