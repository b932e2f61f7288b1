set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/r2z_gpu_info.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_gpu.log 2>&1; tail -3 gpurun_out/r2z_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2z_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2z_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-sparse --configs none > gpurun_out/r2z_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
