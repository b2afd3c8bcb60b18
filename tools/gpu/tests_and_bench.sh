mkdir -p gpurun_out
TAG=${1:-q}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2_gputest_$TAG.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2_gputest_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2_bench_$TAG.json 2> gpurun_out/r2_bench_$TAG.err; echo "bench rc $?"; tail -3 gpurun_out/r2_bench_$TAG.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_$TAG.json').read().strip().splitlines()[-1])
print('step', d['ms_per_step'], 'value %.3e'%d['value'], 'e2e', d['e2e']['ms_per_step'], '%.3e'%d['e2e']['value'], d['roofline']['all_kernels_ms'], 'launches', d['gpu_launches'])
P
