# usage: bash tools/gpu/ab_variants.sh "name ..." workload...   — A/B timing of tools/_build/lib<name>.so builds against the shipped library
mkdir -p gpurun_out
NAMES=$1; shift
for v in $NAMES; do python tools/variant.py tools/_build/lib$v.so "$@" 2>&1 | tail -n $#; done > gpurun_out/r2_ab.log 2>&1
python tools/variant.py biorseo_b200/lib/libbiorseo_b200.so "$@" >> gpurun_out/r2_ab.log 2>&1
cat gpurun_out/r2_ab.log
