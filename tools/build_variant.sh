#!/bin/bash
# build_variant.sh NAME "-DFOO=1 ..." : a second libfcwdm build (tools/_bin/libfcwdm_NAME.so) for same-box A/B runs via FCWDM_LIB_PATH
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/tools/_bin/obj_$NAME
mkdir -p $OUT
for f in $ROOT/fast-cwdm_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -I $ROOT/include -c $f -o $OUT/$b.o &
done
wait
nvcc -shared -o $ROOT/tools/_bin/libfcwdm_$NAME.so $OUT/*.o -lcudart -Xlinker --no-as-needed
rm -rf $OUT
echo built tools/_bin/libfcwdm_$NAME.so
