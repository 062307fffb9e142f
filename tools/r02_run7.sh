mkdir -p gpurun_out
for two in 0 1; do
  S2V_GEMM_2CTA=$two timeout 600 python tools/clock_trace.py > gpurun_out/r02_clock_trace_2cta$two.json 2> gpurun_out/r02_clock_trace_2cta$two.err
  echo "trace 2cta=$two rc=$?"; tail -2 gpurun_out/r02_clock_trace_2cta$two.err; cat gpurun_out/r02_clock_trace_2cta$two.json
  S2V_GEMM_2CTA=$two timeout 600 python bench.py --steps 6 --no-sub-runs --no-cpu-baseline --no-library-baseline --no-e2e > gpurun_out/r02_bench_e_2cta$two.json 2> gpurun_out/r02_bench_e_2cta$two.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_e_2cta$two.json').read().strip().splitlines()[-1]); print('2cta=$two', d['ms_per_step'], d['energy'], d['clocks'], {k: v['avg_ms'] for k, v in d['kernels'].items()})"
done
