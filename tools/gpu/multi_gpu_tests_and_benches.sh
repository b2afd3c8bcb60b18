mkdir -p gpurun_out
N=${1:-2}
echo skipping tests
for WL in c4 c4_400k c4b_all; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --workload $WL > gpurun_out/r2_bench_n${N}_strong_$WL.json 2> gpurun_out/r2_bench_n${N}_strong_$WL.err; echo "bench $WL rc $?"; tail -3 gpurun_out/r2_bench_n${N}_strong_$WL.err | cut -c1-300
done
