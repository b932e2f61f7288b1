set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_gpu.log 2>&1; tail -3 gpurun_out/r2u_gpu.log
python tools/roofline_all.py --only bernoulli,negbin,negbin_real,poisson,laplace,studentt,hetero,cat_bij_K100,cat_K100 --clocks > gpurun_out/r2u_roofline.txt 2>&1; tail -4 gpurun_out/r2u_roofline.txt
python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; tail -c 600 gpurun_out/r2u_bench.err
python bench.py --impl reference > gpurun_out/r2u_ref.json 2> gpurun_out/r2u_ref.err
python -c "
import json
d=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['parts'], d['roofline']['frac'], d['e2e']['value'])
for k,v in d['configs'].items(): print(k, v['ms_per_step'], v['cavi']['ms'], v['cavi']['roofline']['frac'], v['gibbs']['ms'])
"
