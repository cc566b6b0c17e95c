#!/bin/bash
# Developer helper: build a variant of libsvfm.so with extra -D flags for A/B runs on the GPU box.
#   tools/build_variant.sh NAME "-DSVFM_ROUND_LB=4" [file.cu ...]   (default: inst_p32_v64.cu, the bench's type)
# The variant lands in tools/dev_libs/libsvfm_NAME.so (git-ignored; travels with gpurun); select it with SVFM_LIB_PATH.
set -e
NAME=$1; FLAGS=$2; shift 2
FILES=${@:-inst_p32_v64.cu}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/sview_fmindex_b200/csrc
OUT=$ROOT/tools/dev_libs
mkdir -p $OUT/$NAME
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
OBJS=""
for f in svfm_api.cu builder.cu inst_p32_v32.cu inst_p32_v64.cu inst_p32_v128.cu inst_p64_v32.cu inst_p64_v64.cu inst_p64_v128.cu; do
  o=$SRC/${f%.cu}.o
  for v in $FILES; do
    if [ "$v" == "$f" ]; then
      o=$OUT/$NAME/${f%.cu}.o
      $NVCC -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr $FLAGS -c $SRC/$f -o $o &
    fi
  done
  [ -f "$SRC/$f" ] && OBJS="$OBJS $o"
done
wait
$NVCC $ARCH -shared -o $OUT/libsvfm_$NAME.so $OBJS -lcudart_static -lpthread -ldl -lrt
echo built $OUT/libsvfm_$NAME.so
