#!/bin/bash
# Build an experimental variant of libalp_b200.so into variants/ (git-ignored; travels to the GPU box):
#   tools/build_variant.sh <name> <-D flags...>      -> variants/libalp_b200_<name>.so
# Only the translation units named in UNITS (default: the two encode TUs) are recompiled with the flags.
set -eu
NAME=$1; shift
UNITS=${UNITS:-"alp_k_encode_f64 alp_k_encode_f32 alp_k_encode_f64u alp_k_encode_f32u"}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OBJ=${ALPB200_OBJ_DIR:-/tmp/alpb200_build/obj}; VOBJ=$(dirname $OBJ)/obj_$NAME
mkdir -p $VOBJ $ROOT/variants
for u in $UNITS; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I$ROOT/alp_b200/csrc -I$ROOT/include "$@" -c $ROOT/alp_b200/csrc/$u.cu -o $VOBJ/$u.o &
done
wait
OBJS=""
for o in $OBJ/*.o; do b=$(basename $o); if [ -f $VOBJ/$b ]; then OBJS="$OBJS $VOBJ/$b"; else OBJS="$OBJS $o"; fi; done
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/variants/libalp_b200_$NAME.so $OBJS
echo built variants/libalp_b200_$NAME.so
