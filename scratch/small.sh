run() { python bench.py --workload $1 --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$2', d['config']['workload'], 'ms/step', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['all_kernels_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],4))
"; }
for w in c1 c2; do
run $w "auto"
BSS_DEBUG_GROUP=8 run $w "G=8"
done
run c3 "auto"
