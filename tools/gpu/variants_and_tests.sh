mkdir -p gpurun_out
TAG=${1:-d1}
timeout 600 python tools/window_variants.py c4 c4b_all c2 c1 > gpurun_out/r2_variants_$TAG.jsonl 2> gpurun_out/r2_variants_$TAG.err; echo "variants rc $?"
python - <<P
import json
for l in open('gpurun_out/r2_variants_$TAG.jsonl'):
    d=json.loads(l); print(d['workload'],d['variant'],'step %.4f count %.4f emit %.4f'%(d['step_ms'],d['count_ms'],d['emit_ms']),d['equal_to_first_variant'],d['window_items'],d['launches'])
P
tail -3 gpurun_out/r2_variants_$TAG.err
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2_gputest_$TAG.log 2>&1; echo "pytest rc $?"
tail -12 gpurun_out/r2_gputest_$TAG.log
