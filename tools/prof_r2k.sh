set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cavi.py tests/test_gpu_gibbs.py tests/test_gpu_host.py -m gpu -x -q > gpurun_out/r2k_gpu.log 2>&1; tail -3 gpurun_out/r2k_gpu.log
python tools/roofline_all.py --only bernoulli --clocks > gpurun_out/r2k_roofline.txt 2> gpurun_out/r2k_roofline.err; tail -4 gpurun_out/r2k_roofline.txt
python tools/roofline_rest.py --only bernoulli,poisson,hetero > gpurun_out/r2k_rest.txt 2>&1; tail -20 gpurun_out/r2k_rest.txt
ncu --set full --clock-control none --import-source on -k regex:'pg1_compact_kernel' -s 1 -c 1 -o gpurun_out/prof_pg1_r2k -f python tools/roofline_all.py --only bernoulli --n 100000000 --reps 1 > gpurun_out/ncu_r2k.log 2>&1
timeout 600 python -m pytest tests/test_gpu_pglaw.py -m gpu -x -q -k "ks_at_1e8" > gpurun_out/r2k_ks.log 2>&1; tail -3 gpurun_out/r2k_ks.log
