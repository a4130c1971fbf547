#!/bin/bash
# builds experiment variants of libnnuepack.so into variants/<name>.so: tools/build_variants.sh name "-DFLAG ..." [name flags ...]
cd "$(dirname "$0")/../nnue_data_compress_b200/csrc" || exit 1
mkdir -p ../../variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  make -s -j8 OUT=../../variants/$name.so OBJDIR=build_$name EXTRA="$flags" ../../variants/$name.so 2>&1 | grep -E "error|rror:" 
  grep -A2 "k_walk_runs\|k_emit_chains_verify" build_$name/compress.log build_$name/decompress.log | grep -E "registers|spill" | sed "s/^/$name: /"
done
