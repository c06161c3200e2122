#!/bin/bash
# A/B timing aid: the library with recurrent_cluster.cu taken from another commit (everything else from the working tree).
#   tools/build_kernel_variant.sh <commit> <out.so> [extra nvcc flags]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
cp -r "$ROOT/bernoulli_var_speech_codec_b200/csrc" "$TMP/csrc"
git -C "$ROOT" show "$1:bernoulli_var_speech_codec_b200/csrc/recurrent_cluster.cu" > "$TMP/csrc/recurrent_cluster.cu"
shift; OUT=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I "$ROOT/include" -I "$TMP/csrc" "$@" -shared -o "$OUT" "$TMP"/csrc/*.cu -lcudart
rm -rf "$TMP"
echo "$OUT"
