mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cta_pair or linear_full" > gpurun_out/r02_run4_pair.log 2>&1
echo "pair tests rc=$?"; tail -15 gpurun_out/r02_run4_pair.log
for two in 0 1; do
  S2V_GEMM_2CTA=$two timeout 200 python tools/gemm_raster_probe.py >> gpurun_out/r02_gemm_2cta$two.jsonl 2>&1
  S2V_GEMM_2CTA=$two timeout 200 python tools/power_probe.py gemm_ffn_up torch_matmul_ffn_up >> gpurun_out/r02_power_2cta$two.jsonl 2>&1
  echo "== 2cta=$two"; cat gpurun_out/r02_gemm_2cta$two.jsonl | cut -c1-150; cat gpurun_out/r02_power_2cta$two.jsonl
done
S2V_GEMM_2CTA=1 ITERS=3 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/r02_gemm_2cta1_ncu.csv python tools/gemm_raster_probe.py > /dev/null 2>&1
S2V_GEMM_2CTA=0 ITERS=3 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/r02_gemm_2cta0_ncu.csv python tools/gemm_raster_probe.py > /dev/null 2>&1
for two in 0 1; do
  S2V_GEMM_2CTA=$two timeout 600 python bench.py --steps 5 --no-sub-runs --no-cpu-baseline --no-library-baseline --no-e2e > gpurun_out/r02_bench_2cta$two.json 2> gpurun_out/r02_bench_2cta$two.err
  echo "bench 2cta=$two rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_2cta$two.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernels'])"
done
