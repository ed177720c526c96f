#!/bin/bash
# Experimental build of the library: one source file recompiled with extra flags, linked against the default
# objects.   build_variant.sh NAME FILE.cu "-DFLAG=..."   ->  calibration-normalizing-flows_b200/libcnf_NAME.so
# (select it with CNF_B200_LIB=<path>; timing experiments only).
set -e
NAME=$1; FILE=$2; FLAGS=$3
CSRC="$(cd "$(dirname "$0")/../../calibration-normalizing-flows_b200/csrc" && pwd)"
cd "$CSRC"
make -s all
mkdir -p build_$NAME
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $FLAGS -c $FILE -o build_$NAME/${FILE%.cu}.o
OBJS=""
for o in build/*.o; do
  b=$(basename $o)
  if [ "$b" == "${FILE%.cu}.o" ]; then OBJS="$OBJS build_$NAME/$b"; else OBJS="$OBJS $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libcnf_$NAME.so $OBJS -cudart static
echo "built $(cd .. && pwd)/libcnf_$NAME.so"
