mkdir -p gpurun_out
timeout 300 python tools/two_stream_probe.py serial | tee -a gpurun_out/r02_two_stream.jsonl
for sms in 40 52 64 80 148; do
  S2V_GEMM_SMS=$sms timeout 300 python tools/two_stream_probe.py split | tee -a gpurun_out/r02_two_stream.jsonl
done
timeout 300 python tools/two_stream_probe.py serial | tee -a gpurun_out/r02_two_stream.jsonl
