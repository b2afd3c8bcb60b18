mkdir -p gpurun_out
for v in cta10 cta12; do python tools/variant.py tools/_build/lib$v.so c4 c4b c4_400k 2>&1 | tail -3; done > gpurun_out/r2_regs2.log 2>&1
cat gpurun_out/r2_regs2.log
