#!/bin/bash
# Links tools/hoststub/libbcnn_hoststub.so: the real host runtime objects (bcnn_b200/_build, built
# by `python -m bcnn_b200.build`) with device.cu / optim.cu replaced by device_stub.cpp.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
B=$HERE/../../bcnn_b200/_build
g++ -O1 -fPIC -ffp-contract=off -c $HERE/device_stub.cpp -o $HERE/device_stub.o
OBJS=$(ls $B/*.o | grep -v "/device.cu.o\|/optim.cu.o")
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $HERE/libbcnn_hoststub.so \
    $HERE/device_stub.o $OBJS -Xcompiler -fPIC --cudart=static -Xlinker -Bsymbolic -ldl -lm -lpthread -lrt
echo $HERE/libbcnn_hoststub.so
