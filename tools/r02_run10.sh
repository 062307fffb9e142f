mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vae.py -x -q -m gpu > gpurun_out/r02_run10_vae.log 2>&1; echo "vae tests rc=$?"; tail -3 gpurun_out/r02_run10_vae.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "block or transformer or pipeline or fp16 or graph" > gpurun_out/r02_run10_par.log 2>&1; echo "parity subset rc=$?"; tail -3 gpurun_out/r02_run10_par.log
for two in 0 1; do S2V_GEMM_2CTA=$two timeout 300 python tools/vae_bench.py 2>&1 | grep vae_decode | tee -a gpurun_out/r02_vae_bench.jsonl; done
S2V_ADALN_ROWS=1 timeout 300 python tools/hbm_kernels_bench.py 2>&1 | grep adaln | tee -a gpurun_out/r02_hbm_kernels.jsonl
timeout 300 python tools/hbm_kernels_bench.py 2>&1 | grep kernel | tee -a gpurun_out/r02_hbm_kernels.jsonl
