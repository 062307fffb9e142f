mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_t5.py -x -q -m gpu > gpurun_out/r02_run13_t5.log 2>&1; echo "t5 tests rc=$?"; tail -25 gpurun_out/r02_run13_t5.log
timeout 600 python tools/t5_bench.py 2>&1 | tail -3 | tee gpurun_out/r02_t5_bench.json
