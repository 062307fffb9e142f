mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vae.py -x -q -m gpu > gpurun_out/r02_run11_vae.log 2>&1; echo "vae tests rc=$?"; tail -5 gpurun_out/r02_run11_vae.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cta_pair or linear_full or block_cfg1 or adaln or full_step" > gpurun_out/r02_run11_par.log 2>&1; echo "parity subset rc=$?"; tail -3 gpurun_out/r02_run11_par.log
for f in 0 1; do S2V_VAE_FUSED_GN=$f timeout 300 python tools/vae_bench.py 2>&1 | grep vae_decode | sed "s/^{/{\"fused_gn\": $f, /" | tee -a gpurun_out/r02_vae_bench2.jsonl; done
