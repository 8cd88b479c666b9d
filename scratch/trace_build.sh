#!/bin/bash
# scratch build of the library with the persistent-kernel timeline compiled in
cd /root/repo/mola_lidar_odometry_b200
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -DMLO_TRACE -Xcompiler -fPIC,-O3,-ffp-contract=off,-pthread -shared \
  -I ../include -I csrc -I host -o ../scratch/libmlo_b200_trace.so csrc/mlo_b200.cu host/host_capi.cpp -lcudart
