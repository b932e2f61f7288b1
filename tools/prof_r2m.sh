set -x
mkdir -p gpurun_out
for p in 21 22 42; do
  AUGCUDA_CAT_PIPE=$p timeout 600 python -m pytest tests/test_gpu_cavi.py tests/test_gpu_host.py -m gpu -x -q -k "cat or Cat or CAT" > gpurun_out/r2m_gpu_$p.log 2>&1; tail -2 gpurun_out/r2m_gpu_$p.log
  AUGCUDA_CAT_PIPE=$p python tools/roofline_all.py --only cat_bij_K100,cat_K100 > gpurun_out/r2m_roofline_$p.txt 2> gpurun_out/r2m_roofline_$p.err; tail -3 gpurun_out/r2m_roofline_$p.txt
done
ncu --set full --clock-control none --import-source on -k regex:'cat_row_kernel' -s 0 -c 1 -o gpurun_out/prof_catrow_elbo_r2m -f python tools/roofline_all.py --only cat_bij_K100 --ncat 2000000 --reps 1 > gpurun_out/ncu_r2m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'cat_gibbs_kernel' -s 1 -c 1 -o gpurun_out/prof_catgibbs_r2m -f python tools/roofline_all.py --only cat_bij_K100 --ncat 2000000 --reps 1 > gpurun_out/ncu_r2m2.log 2>&1
