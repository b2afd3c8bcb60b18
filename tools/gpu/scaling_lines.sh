# usage: bash tools/gpu/r2_scale.sh N workload...   — strong-scaling bench lines (with the weak number beside), per-rank table printed
mkdir -p gpurun_out
N=$1; shift
for WL in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --steps 20 --warmup 5 --workload $WL --no-cpu-baseline --no-extra > gpurun_out/r2_bench_n1_$WL.json 2> gpurun_out/r2_bench_n1_$WL.err; echo "bench $WL rc $?"
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --workload $WL --no-cpu-baseline > gpurun_out/r2_bench_n${N}_strong_$WL.json 2> gpurun_out/r2_bench_n${N}_strong_$WL.err; echo "bench $WL rc $?"
  fi
  tail -2 gpurun_out/r2_bench_n${N}_*$WL.err | cut -c1-300
done
