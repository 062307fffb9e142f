mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants or lags" > gpurun_out/r02_run2_variants.log 2>&1
echo "variants rc=$?"; tail -3 gpurun_out/r02_run2_variants.log
timeout 600 python tools/attn_ab.py > gpurun_out/r02_attn_ab2.jsonl 2> gpurun_out/r02_attn_ab2.err
echo "attn_ab rc=$?"
timeout 900 python tools/instep_ab.py > gpurun_out/r02_instep_ab2.jsonl 2> gpurun_out/r02_instep_ab2.err
echo "instep rc=$?"
timeout 600 python tools/power_probe.py > gpurun_out/r02_power_probe.jsonl 2> gpurun_out/r02_power_probe.err
echo "power rc=$?"
for h in 1 2; do for g in 16 32; do
  S2V_GEMM_L2_HINT=$h S2V_GEMM_GROUP_M=$g timeout 200 python tools/gemm_raster_probe.py > gpurun_out/r02_gemm_hint${h}_g$g.jsonl 2>&1
  S2V_GEMM_L2_HINT=$h S2V_GEMM_GROUP_M=$g ITERS=3 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/r02_gemm_hint${h}_ncu_g$g.csv python tools/gemm_raster_probe.py > /dev/null 2>&1
done; done
echo "raster done"
