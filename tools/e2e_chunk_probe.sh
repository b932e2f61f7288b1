# usage (on the GPU box): bash tools/e2e_chunk_probe.sh   — e2e leg of bench.py for several host chunk sizes
python tools/pcie_probe.py
for lg in 20 21 22 23 24; do
  AUGCUDA_HOST_CHUNK_LOG2=$lg timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu --no-sparse 2>/dev/null > gpurun_out/e2e_lg$lg.json
  python -c "import json; d=json.load(open('gpurun_out/e2e_lg$lg.json')); print('chunk_log2', $lg, 'ms', d['e2e']['ms_per_step'], 'obs/s', d['e2e']['value'])"
done
