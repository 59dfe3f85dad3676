#!/bin/bash
# rebuild libgpsiq.so in-tree (same flags as __graft_entry__.build)
cd "$(dirname "$0")/../pluto_gps_sim_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false \
  -Xcompiler -fPIC -shared "$@" -o ../libgpsiq.so gpsiq.cu
