#!/bin/bash
# usage: scripts/build_variant.sh <name> <extra nvcc flags...>  -> libvadc_<name>.so with cluster_bwd_tc2.cu rebuilt under the flags
# (kernel-variant experiments: VADC_LIB_PATH=<that file> selects it at run time; SRC=<file.cu> picks another source)
set -e
D=$(dirname "$0")/../video-anomaly-detection-guided-by-clustering-learning_b200
N=$1; shift
SRC=${SRC:-cluster_bwd_tc2.cu}
B=${SRC%.cu}
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 -cudart static "$@" -c $D/csrc/$SRC -o $D/build/variant_${B}_$N.obj
OBJS=$(ls $D/build/*.o | grep -v "/$B.o" | tr '\n' ' ')
nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o $D/libvadc_$N.so $OBJS $D/build/variant_${B}_$N.obj -ldl -lpthread
echo built $D/libvadc_$N.so
