mkdir -p gpurun_out
DRAW_SWEEP=1 timeout 600 python tools/window_variants.py c4 c4b > gpurun_out/r2_variants_v9.jsonl 2> gpurun_out/r2_variants_v9.err; echo "variants rc $?"
tail -3 gpurun_out/r2_variants_v9.err
