# A/B of two builds of the library on cases of tools/roofline_all.py:  bash tools/ab.sh CASES variants/libaugcuda_X.so [reps]
CASE=$1; ALT=$2; R=${3:-10}
for i in 1 2; do
  echo "stock:"; python tools/roofline_all.py --only $CASE --reps $R 2>&1 | grep "^| [a-z]"
  echo "alt $ALT:"; AUGCUDA_LIB=$ALT python tools/roofline_all.py --only $CASE --reps $R 2>&1 | grep "^| [a-z]"
done
