#!/bin/bash
# build a tuning variant of the library: profiles/mkvariant.sh name "-DCAMCAL_TL_F32=32 ..."
# -> profiles/variants/lib_<name>.so   (only rectify.cu is recompiled; other objects are reused)
set -e
cd "$(dirname "$0")/../cameracalibrations_b200/csrc"
mkdir -p ../../profiles/variants _build
make -s -j8 >/dev/null
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off"
$NV $2 -c rectify.cu -o _build/rectify_$1.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../../profiles/variants/lib_$1.so _build/abi.o _build/chain_host.o _build/pointmap.o _build/residual.o _build/lm.o _build/comm.o _build/init.o _build/ingest.o _build/rectify_$1.o -ldl
echo built lib_$1.so
