mkdir -p gpurun_out
for two in 0 1; do
  S2V_GEMM_2CTA=$two timeout 600 python tools/clock_trace.py > gpurun_out/r02_clock_trace_2cta$two.json 2> gpurun_out/r02_clock_trace_2cta$two.err
  echo "trace 2cta=$two rc=$?"; tail -2 gpurun_out/r02_clock_trace_2cta$two.err; cat gpurun_out/r02_clock_trace_2cta$two.json
done
REPS=1 timeout 900 python tools/instep_ab.py hi_mc=3:1:200 r1_lo=0:1:200 > gpurun_out/r02_instep_ab3.jsonl 2> gpurun_out/r02_instep_ab3.err
echo "instep rc=$?"; cat gpurun_out/r02_instep_ab3.jsonl
rm -f gpurun_out/r02_parity.jsonl
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputest.log 2>&1
echo "gputest rc=$?"; tail -4 gpurun_out/r02_gputest.log
