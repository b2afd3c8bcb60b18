mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2_gputest_b3.log 2>&1; echo "pytest rc $?"
tail -12 gpurun_out/r2_gputest_b3.log
timeout 600 python tools/window_variants.py c4 c4b c4a c4_desc c2 > gpurun_out/r2_variants_b3.jsonl 2> gpurun_out/r2_variants_b3.err; echo "variants rc $?"
cut -c1-330 gpurun_out/r2_variants_b3.jsonl; tail -3 gpurun_out/r2_variants_b3.err
