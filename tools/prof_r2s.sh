set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gibbs.py tests/test_gpu_pglaw.py tests/test_gpu_host.py tests/test_gpu_testutils.py -m gpu -x -q > gpurun_out/r2s_gpu.log 2>&1; tail -3 gpurun_out/r2s_gpu.log
python tools/roofline_all.py --only cat_bij_K100,cat_K100 > gpurun_out/r2s_roofline.txt 2>&1; tail -5 gpurun_out/r2s_roofline.txt

timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_path.py > gpurun_out/r2s_race.log 2>&1; echo "race rc=$?"; tail -2 gpurun_out/r2s_race.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_path.py > gpurun_out/r2s_mem.log 2>&1; echo "mem rc=$?"; tail -2 gpurun_out/r2s_mem.log
