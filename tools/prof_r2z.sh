mkdir -p gpurun_out
bash tools/ab.sh bernoulli 20 variants/libaugcuda_sup4.so variants/libaugcuda_sup16.so > gpurun_out/ab_super.txt 2>&1
AUGCUDA_LIB=variants/libaugcuda_sup4.so timeout 300 python -m pytest tests/test_gpu_gibbs.py -m gpu -x -q -k "not categorical" 2>&1 | tail -2 >> gpurun_out/ab_super.txt
