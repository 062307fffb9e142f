mkdir -p gpurun_out
for two in 1 0; do for sms in 148 132 116 100; do
  S2V_GEMM_SMS=$sms S2V_GEMM_2CTA=$two timeout 600 python bench.py --steps 6 --no-sub-runs --no-cpu-baseline --no-library-baseline --no-e2e > gpurun_out/r02_bench_sms.json 2> gpurun_out/r02_bench_sms.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_sms.json').read().strip().splitlines()[-1]); print(json.dumps({'2cta': $two, 'gemm_sms': $sms, 'ms_per_step': d['ms_per_step'], 'energy': d['energy'], 'sm_mhz': d['clocks']['sm_mhz'], 'kernels': {k: v['avg_ms'] for k, v in d['kernels'].items()}}))" | tee -a gpurun_out/r02_gemm_sms_sweep.jsonl
done; done
