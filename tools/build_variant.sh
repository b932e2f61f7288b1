#!/bin/bash
# A/B builds: tools/build_variant.sh NAME "EXTRA nvcc flags" file.cu [file.cu ...]
# recompiles the named sources with the extra flags into variants/NAME/ and links variants/libaugcuda_NAME.so from them plus
# the stock objects; select it with AUGCUDA_LIB=variants/libaugcuda_NAME.so (variants/ is git-ignored but travels to the GPU box)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/augmentedgplikelihoods.jl_b200/csrc
NAME=$1; EXTRA=$2; shift 2
mkdir -p $ROOT/variants/$NAME
OBJS=""
for f in aug_ctx aug_cavi aug_cat aug_gibbs aug_loglik aug_host aug_next aug_sparse; do
  if [[ " $* " == *" $f.cu "* ]]; then
    FMAD=true; [[ $f == aug_cavi || $f == aug_cat ]] && FMAD=false
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ --fmad=$FMAD $EXTRA -c $C/$f.cu -o $ROOT/variants/$NAME/$f.o
    OBJS="$OBJS $ROOT/variants/$NAME/$f.o"
  else
    OBJS="$OBJS $C/$f.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/variants/libaugcuda_$NAME.so $OBJS -cudart static -ldl
echo variants/libaugcuda_$NAME.so
