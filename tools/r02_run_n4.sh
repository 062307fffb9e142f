mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544"
timeout 1200 $TR bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
echo "n4 rc=$?"; tail -3 gpurun_out/r02_bench_n4.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step']); print(json.dumps(d.get('cfg_sharded'))); print(json.dumps(d.get('cfg4'))); print(json.dumps(d.get('video_e2e'))[:500])"
