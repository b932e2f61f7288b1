# A/B of builds of the library on cases of tools/roofline_all.py:  bash tools/ab.sh CASES REPS variants/libaugcuda_X.so [more .so ...]
CASE=$1; R=$2; shift 2
for i in 1 2; do
  echo "stock:"; python tools/roofline_all.py --only $CASE --reps $R 2>&1 | grep "^| [a-z]"
  for ALT in "$@"; do
    echo "alt $ALT:"; AUGCUDA_LIB=$ALT python tools/roofline_all.py --only $CASE --reps $R 2>&1 | grep "^| [a-z]"
  done
done
