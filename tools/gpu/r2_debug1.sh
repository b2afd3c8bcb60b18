mkdir -p gpurun_out
for v in "" nilevel assert; do
echo "== lib '$v'"; BSS_LIB=$v python tools/gpu/debug_synth.py 3 230 jsonfolder 2>&1 | tail -1 | cut -c1-150
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin_cpp.py -m gpu -q -p no:cacheprovider --maxfail=10 -k "golden or window or matches_interpreter or dropin" > gpurun_out/r2_gputest_v4.log 2>&1; echo "pytest rc $?"
tail -15 gpurun_out/r2_gputest_v4.log | cut -c1-300
