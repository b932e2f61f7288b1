set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'cat_gibbs_kernel' -s 1 -c 1 -o gpurun_out/prof_catgibbs_r2z -f python tools/roofline_all.py --only cat_bij_K100 --ncat 2000000 --reps 1 > gpurun_out/ncu_r2z.log 2>&1
