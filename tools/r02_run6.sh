mkdir -p gpurun_out
for two in 0 1; do
  S2V_GEMM_2CTA=$two timeout 600 python tools/clock_trace.py > gpurun_out/r02_clock_trace_2cta$two.json 2> gpurun_out/r02_clock_trace_2cta$two.err
  echo "trace 2cta=$two rc=$?"; tail -2 gpurun_out/r02_clock_trace_2cta$two.err; cat gpurun_out/r02_clock_trace_2cta$two.json
done
