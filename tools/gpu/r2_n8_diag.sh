mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --workload c4 --no-cpu-baseline > gpurun_out/r2_bench_n${N}_diag_c4.json 2> gpurun_out/r2_bench_n${N}_diag_c4.err; echo "bench rc $?"; tail -3 gpurun_out/r2_bench_n${N}_diag_c4.err | cut -c1-300
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_n${N}_diag_c4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['limiter'])
for r in d['per_rank']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items()})
print('weak', d['weak_scaling']['ms_per_step'])
for r in d['weak_scaling']['per_rank']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items()})
P
