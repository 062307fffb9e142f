mkdir -p gpurun_out
for f in 0 1; do S2V_VAE_FUSED_GN=$f timeout 300 python tools/vae_bench.py 2>&1 | grep vae_decode | sed "s/^{/{\"fused_gn\": $f, /" | tee -a gpurun_out/r02_vae_bench3.jsonl; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file gpurun_out/r02_launches_vae.csv python tools/vae_bench.py > /dev/null 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r02_launches_vae.csv
