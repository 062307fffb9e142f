mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check.log 2>&1
echo "check rc=$?"; tail -3 gpurun_out/r02_multi_gpu_check.log
timeout 300 $TR bench.py --gpus 2 --workload tiny --steps 3 > gpurun_out/r02_bench_tiny_n2.json 2> gpurun_out/r02_bench_tiny_n2.err
echo "tiny n2 rc=$?"; tail -3 gpurun_out/r02_bench_tiny_n2.err; head -c 300 gpurun_out/r02_bench_tiny_n2.json
timeout 900 $TR bench.py --gpus 2 --steps 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo "n2 rc=$?"; tail -3 gpurun_out/r02_bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], json.dumps(d.get('cfg_sharded')), json.dumps(d.get('video_e2e'))[:600])"
