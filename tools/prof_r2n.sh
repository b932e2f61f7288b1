set -x
mkdir -p gpurun_out
for p in 161 162 321 322 641; do
  AUGCUDA_CAT_PIPE=$p timeout 600 python -m pytest tests/test_gpu_cavi.py tests/test_gpu_host.py -m gpu -x -q -k "cat or Cat or CAT" > gpurun_out/r2n_gpu_$p.log 2>&1; tail -2 gpurun_out/r2n_gpu_$p.log
  AUGCUDA_CAT_PIPE=$p python tools/roofline_all.py --only cat_bij_K100,cat_K100 > gpurun_out/r2n_roofline_$p.txt 2> gpurun_out/r2n_roofline_$p.err; tail -2 gpurun_out/r2n_roofline_$p.txt
done
python tools/roofline_all.py --only negbin,poisson,hetero > gpurun_out/r2n_roofline_pgb.txt 2>&1; tail -4 gpurun_out/r2n_roofline_pgb.txt
