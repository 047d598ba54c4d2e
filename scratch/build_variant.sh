#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. ..." : libsem_b200_NAME.so with the P=8 marching TU recompiled with extra flags (A/B runs)
set -e
cd "$(dirname "$0")/../sem_b200/csrc"
NAME=$1; FLAGS=$2
mkdir -p build/var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $FLAGS -DSEM_P=8 -c sem_march_inst.cu -o build/var/p8_$NAME.o 2> build/var/p8_$NAME.log
OBJS=$(ls build/*.o | grep -v "sem_march_p8.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/libsem_b200_$NAME.so $OBJS build/var/p8_$NAME.o -ldl
grep -A2 "march3" build/var/p8_$NAME.log | grep -E "registers" | tr '\n' ' '; echo
