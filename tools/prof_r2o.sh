set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cavi.py tests/test_gpu_host.py tests/test_gpu_optimality.py tests/test_gpu_testutils.py tests/test_quirks.py -m gpu -x -q > gpurun_out/r2o_gpu.log 2>&1; tail -3 gpurun_out/r2o_gpu.log
for p in 322 321 161; do
  AUGCUDA_CAT_PIPE=$p python tools/roofline_all.py --only cat_bij_K100 > gpurun_out/r2o_roofline_$p.txt 2> gpurun_out/r2o_roofline_$p.err; tail -1 gpurun_out/r2o_roofline_$p.txt
done
