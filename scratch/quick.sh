#!/bin/bash
# quick GPU iteration: fuzz + goldens + the two c4 timings
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random or golden or overlapping" 2>&1 | tail -3
for w in c4 c4b_all; do
python bench.py --workload $w --steps 10 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['config']['workload'], 'ms/step', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['all_kernels_ms'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))
"
done
