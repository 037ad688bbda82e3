#!/bin/bash
# Builds tools/cabi_check (hardware check of the C ABI without Python) against the in-tree library.
set -e
cd "$(dirname "$0")/.."
PKG=generalized-class-discovery-for-lidar-semantic-segmentation_b200
make -C $PKG/csrc -j8 > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Iinclude -o tools/cabi_check tools/cabi_check.cu \
     -L$PKG/gcdlss_b200 -lgcdlss_sm100a -Xlinker -rpath -Xlinker "\$ORIGIN/../$PKG/gcdlss_b200"
echo built tools/cabi_check
