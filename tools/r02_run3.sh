mkdir -p gpurun_out
timeout 300 python bench.py --workload tiny --steps 3 > gpurun_out/r02_bench_tiny.json 2> gpurun_out/r02_bench_tiny.err
echo "tiny bench rc=$?"; tail -3 gpurun_out/r02_bench_tiny.err; head -c 600 gpurun_out/r02_bench_tiny.json
rm -f gpurun_out/r02_parity.jsonl
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputest.log 2>&1
echo "gputest rc=$?"; tail -4 gpurun_out/r02_gputest.log
timeout 1200 python bench.py --steps 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 2 -f -o gpurun_out/r02_attn python tools/profile_kernels.py attn 3 > gpurun_out/r02_ncu_attn.log 2>&1
echo "ncu rc=$?"
for spec in "qkv 37" "qkv 40" "ffn_up 37" "ffn_up 40" "ffn_down 8" "ffn_down 12" "ffn_down 24" "out 12" "out 24"; do
  set -- $spec
  S2V_GEMM_GROUP_M=$2 ITERS=3 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/r02_gemm_raster3_$1_g$2.csv python tools/gemm_raster_probe.py $1 > /dev/null 2>&1
  S2V_GEMM_GROUP_M=$2 timeout 200 python tools/gemm_raster_probe.py $1 >> gpurun_out/r02_gemm_raster3.jsonl 2>&1
done
echo "raster done"
