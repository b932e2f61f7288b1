# usage: bash tools/run_multi.sh N   (inside gpurun --gpus N)
N=${1:-2}
nvidia-smi -L > gpurun_out/multi_gpus.txt
nvidia-smi topo -m >> gpurun_out/multi_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
for coll in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 200 --warmup 5 --collective $coll --no-cpu > gpurun_out/bench_n${N}_${coll}.json 2> gpurun_out/bench_n${N}_${coll}.err
  echo "rc=$?"; cat gpurun_out/bench_n${N}_${coll}.json; tail -3 gpurun_out/bench_n${N}_${coll}.err
done
