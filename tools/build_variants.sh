#!/bin/bash
# Dev tool: builds timing-stamped variants of the fused kernel (compile-time TCV_* knobs) as libsdfb200_v<name>.so for same-box A/B runs
# with tools/tc_timing.py <precision> <log2T> <table dtype> <fused|unfused> <variant name>.   usage: tools/build_variants.sh name:"-DTCV_X=1 ..." ...
set -e
cd "$(dirname "$0")/../sdfstudio_b200"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
OBJS=$(ls build/*.o | grep -v "field_tc_p2_torch\|field_tc_v\|field_tc_host_v\|build/field_tc.o\|tc_test" )
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  /usr/local/cuda/bin/nvcc $FLAGS -DSDFB200_TC_TIMING $defs -c csrc/field_tc_p2_torch.cu -o build/field_tc_v$name.o &
  /usr/local/cuda/bin/nvcc $FLAGS $defs -c csrc/field_tc.cu -o build/field_tc_host_v$name.o &
done
wait
for spec in "$@"; do
  name="${spec%%:*}"
  /usr/local/cuda/bin/nvcc -shared -o libsdfb200_v$name.so $OBJS build/field_tc_v$name.o build/field_tc_host_v$name.o -lcudart
  echo built libsdfb200_v$name.so
done
