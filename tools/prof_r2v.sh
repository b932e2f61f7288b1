set -x
mkdir -p gpurun_out
for k in negbin poisson hetero; do
ncu --set full --clock-control none --import-source on -k regex:'pgb_kernel' -s 1 -c 1 -o gpurun_out/prof_${k}_r2v -f python tools/roofline_all.py --only $k --n 20000000 --reps 1 > gpurun_out/ncu_r2v_$k.log 2>&1
done
