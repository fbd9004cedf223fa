#!/bin/bash
# tools/build_tune.sh NAME "-DDCRF_TUNE_X=.. ..."  -> wsss_analysis_b200/csrc/tune/libdcrf_NAME.so
# (a tuning build of the library with extra -D flags; select it with DCRF_B200_LIB=<path>)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
OBJ=/tmp/dcrf_tune_$NAME
mkdir -p $OBJ $ROOT/wsss_analysis_b200/csrc/tune
cd $ROOT/wsss_analysis_b200/csrc
for f in api lattice_build filter primitives confusion resize collective; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o $OBJ/$f.o $f.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tune/libdcrf_$NAME.so $OBJ/*.o -ldl
echo built tune/libdcrf_$NAME.so
