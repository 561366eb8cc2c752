#!/bin/bash
# builds tools/microbench (per-SM throughput of the instruction classes the kernels are made of)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/microbench tools/microbench.cu
