set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gibbs.py tests/test_gpu_host.py -m gpu -x -q -k "categorical or cat" 2>&1 | tail -3
python tools/roofline_all.py --only cat_bij_K100 2>&1 | tail -1
python tools/roofline_all.py --only cat_bij_K100 2>&1 | tail -1
