set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gibbs.py tests/test_gpu_host.py -m gpu -x -q -k "categorical or cat" 2>&1 | tail -5
python tools/roofline_all.py --only cat_bij_K100,cat_K100 2>&1 | tail -3
SAN_CATGIBBS=40000 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_path.py > gpurun_out/sanitize_catgibbs_racecheck_r2z.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitize_catgibbs_racecheck_r2z.log
SAN_CATGIBBS=40000 timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_path.py > gpurun_out/sanitize_catgibbs_memcheck_r2z.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_catgibbs_memcheck_r2z.log
