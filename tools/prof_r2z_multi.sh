set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2z_gpus_n$N.txt
if [ -z "$SKIPTESTS" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2z_multi_n$N.log 2>&1; tail -3 gpurun_out/r2z_multi_n$N.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/r2z_bench_n$N.json 2> gpurun_out/r2z_bench_n$N.err
echo rc=$?; tail -c 400 gpurun_out/r2z_bench_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2z_bench_n$N.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],d['parts'])
print('e2e',d['e2e']['value'], 'checks',json.dumps(d.get('checks'))[:600])
print('strong',json.dumps(d.get('strong_scaling'))[:800])
for k,v in d['configs'].items(): print(k, v['ms_per_step'], v.get('obs_per_s'))
print('sparse',json.dumps(d.get('sparse_sweep',{}).get('library_composition'))[:400])
P
