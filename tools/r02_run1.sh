mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_run1_gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants or lags" > gpurun_out/r02_run1_variants.log 2>&1
echo "variants rc=$?"
timeout 600 python tools/attn_ab.py > gpurun_out/r02_attn_ab.jsonl 2> gpurun_out/r02_attn_ab.err
echo "attn_ab rc=$?"
timeout 900 python tools/instep_ab.py > gpurun_out/r02_instep_ab.jsonl 2> gpurun_out/r02_instep_ab.err
echo "instep rc=$?"
for g in 16 32 48 64 96; do
  S2V_GEMM_GROUP_M=$g timeout 200 python tools/gemm_raster_probe.py > gpurun_out/r02_gemm_raster_g$g.jsonl 2>&1
  S2V_GEMM_GROUP_M=$g ITERS=3 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/r02_gemm_raster_ncu_g$g.csv python tools/gemm_raster_probe.py > /dev/null 2>&1
done
echo "raster done"
rm -f gpurun_out/r02_parity.jsonl
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputest.log 2>&1
echo "gputest rc=$?"
tail -5 gpurun_out/r02_gputest.log
