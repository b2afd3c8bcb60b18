#!/bin/bash
run() { python bench.py --workload $1 --steps 10 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$2', d['config']['workload'], 'ms/step', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['all_kernels_ms'].items()})
"; }
for tm in 1024 4000000000; do
BSS_DEBUG_TABMIN=$tm run c4b_all "tabmin=$tm"
BSS_DEBUG_TABMIN=$tm run c4 "tabmin=$tm"
done
