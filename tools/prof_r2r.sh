set -x
mkdir -p gpurun_out
python tools/roofline_all.py --only bernoulli 2>&1 | tail -1
for v in pg1ilp2 pg1ilp3; do
  AUGCUDA_LIB=variants/libaugcuda_$v.so python tools/roofline_all.py --only bernoulli 2>&1 | tail -1
  AUGCUDA_LIB=variants/libaugcuda_$v.so timeout 600 python -m pytest tests/test_gpu_gibbs.py -m gpu -x -q -k "pg_sampler or shard or init_aux" 2>&1 | tail -1
done
