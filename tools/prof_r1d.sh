set -x
python tools/roofline_all.py --clocks > gpurun_out/roofline_all_r1d.txt 2> gpurun_out/roofline_all_r1d.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'cavi_tma_kernel|pg1_compact_kernel' -s 6 -c 2 -o gpurun_out/prof_bench_r1d -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'cat_tma_kernel|cat_gibbs_kernel' -s 4 -c 3 -o gpurun_out/prof_cat_r1d -f python tools/roofline_all.py --only cat_bij_K100 --ncat 2000000 --reps 1 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'aux_sample_kernel' -s 1 -c 1 -o gpurun_out/prof_gibbs_negbin_r1d -f python tools/roofline_all.py --only negbin --n 20000000 --reps 1 > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'aux_sample_kernel' -s 1 -c 1 -o gpurun_out/prof_gibbs_poisson_r1d -f python tools/roofline_all.py --only poisson --n 20000000 --reps 1 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out
tail -12 gpurun_out/roofline_all_r1d.txt
