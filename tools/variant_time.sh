#!/bin/bash
# GPU box: K1 and chain-decoder times of experiment builds: tools/variant_time.sh name ... ("main" = the in-tree library)
N=${N:-100000000}
for v in "$@"; do
  if [ "$v" = main ]; then unset NNP_LIB; else export NNP_LIB=$PWD/variants/$v.so; fi
  c=$(python tools/cmp_time.py $N 100 2>&1 | grep -E "chain-walk|identical" | tr '\n' ' ')
  d=$(python tools/dec_time.py $N 2>&1 | grep -E "decompress rc|re-encode" | tr '\n' ' ')
  echo "$v | $c | $d"
done
