#!/bin/bash
# builds the stand-alone design micro-benchmarks (run them on the GPU box with tools/ubench/run.sh)
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false hist_bench.cu -o hist_bench
