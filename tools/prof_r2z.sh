set -x
mkdir -p gpurun_out
SAN_CATGIBBS=40000 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_path.py > gpurun_out/sanitize_catgibbs_racecheck_r2z.log 2>&1; echo "racecheck rc=$?"; tail -12 gpurun_out/sanitize_catgibbs_racecheck_r2z.log
timeout 600 python -m pytest tests/test_gpu_gibbs.py -m gpu -x -q -k categorical 2>&1 | tail -3
python tools/roofline_all.py --only cat_bij_K100 2>&1 | tail -2
