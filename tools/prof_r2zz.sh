# last GPU call of round 2: ncu --set full of the two kernels of the headline step in their FINAL form
# (pg1_compact_kernel 640 threads / 96 registers, cavi_tma_kernel<BERNOULLI, ELBO>), N = 1e8
mkdir -p gpurun_out
timeout 110 ncu --set full --clock-control none --import-source on -k regex:'pg1_compact_kernel' -s 2 -c 1 -o gpurun_out/prof_pg1_r2zz -f python tools/roofline_all.py --only bernoulli --reps 1 > gpurun_out/ncu_r2zz_pg1.log 2>&1; echo "pg1 rc=$?"
timeout 70 ncu --set full --clock-control none --import-source on -k regex:'cavi_tma_kernel' -s 2 -c 1 -o gpurun_out/prof_cavi_r2zz -f python tools/roofline_all.py --only bernoulli --reps 1 > gpurun_out/ncu_r2zz_cavi.log 2>&1; echo "cavi rc=$?"
