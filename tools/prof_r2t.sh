set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'pgb_kernel' -s 1 -c 1 -o gpurun_out/prof_hetero_r2t -f python tools/roofline_all.py --only hetero --n 20000000 --reps 1 > gpurun_out/ncu_r2t.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pgb_kernel' -s 1 -c 1 -o gpurun_out/prof_poisson_r2t -f python tools/roofline_all.py --only poisson --n 20000000 --reps 1 > gpurun_out/ncu_r2t2.log 2>&1
