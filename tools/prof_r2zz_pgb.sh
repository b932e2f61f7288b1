# ncu --set full of pgb_kernel<NEGBIN | POISSON | HETERO> in the final round-2 CTA shape (one 768-thread CTA per SM), 2e7 observations
mkdir -p gpurun_out
for k in negbin poisson hetero; do
timeout 45 ncu --set full --clock-control none --import-source on -k regex:'pgb_kernel' -s 1 -c 1 -o gpurun_out/prof_${k}_r2zz -f python tools/roofline_all.py --only $k --n 20000000 --reps 1 > gpurun_out/ncu_r2zz_$k.log 2>&1; echo "$k rc=$?"
done
