#!/bin/sh
# A/B builds of the engine library for one-call comparisons on the GPU box (bench.py picks the library named by
# NLZM_MF_LIB). Usage: sh tools/ab/build_variants.sh ; then e.g.
#   gpurun -- 'python bench.py --c3 0 --compress 0 > gpurun_out/a.json; \
#              NLZM_MF_LIB=$PWD/tools/ab/libnlzm_mf_lsu.so python bench.py --c3 0 --compress 0 > gpurun_out/b.json'
cd "$(dirname "$0")/../../nlzm_b200/csrc" || exit 1
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared"
nvcc $FLAGS -DNLZM_MT_TMA=0 -o ../../tools/ab/libnlzm_mf_lsu.so engine.cu     # merge tiles staged by 16-byte LSU loads instead of TMA
