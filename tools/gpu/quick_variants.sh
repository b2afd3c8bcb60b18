mkdir -p gpurun_out
TAG=${1:-v}; shift
timeout 600 python tools/window_variants.py "$@" > gpurun_out/r2_variants_$TAG.jsonl 2> gpurun_out/r2_variants_$TAG.err; echo "variants rc $?"
python - <<P
import json
for l in open('gpurun_out/r2_variants_$TAG.jsonl'):
    d=json.loads(l); print(d['workload'],d['variant'],'step %.4f count %.4f emit %.4f'%(d['step_ms'],d['count_ms'],d['emit_ms']),d['equal_to_first_variant'],d['window_items'],d['launches'])
P
tail -3 gpurun_out/r2_variants_$TAG.err
