set -x
mkdir -p gpurun_out
python tools/roofline_all.py --only bernoulli,cat_bij_K100 > gpurun_out/r2l_roofline.txt 2> gpurun_out/r2l_roofline.err; tail -5 gpurun_out/r2l_roofline.txt
AUGCUDA_LIB=variants/libaugcuda_pg1mb4.so python tools/roofline_all.py --only bernoulli > gpurun_out/r2l_roofline_mb4.txt 2>&1; tail -2 gpurun_out/r2l_roofline_mb4.txt
ncu --set full --clock-control none --import-source on -k regex:'cat_row_kernel|cat_gibbs_kernel' -s 3 -c 3 -o gpurun_out/prof_cat_r2l -f python tools/roofline_all.py --only cat_bij_K100 --ncat 2000000 --reps 1 > gpurun_out/ncu_r2l.log 2>&1
