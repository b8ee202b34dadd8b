#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. ..."  -> <pkg>/libqdm_NAME.so with qdm_gemm_w4ts.cu compiled with the extra flags (A/B timing)
set -e
cd "$(dirname "$0")/../quantization---diffusion-models_b200/csrc"
name=$1; shift
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c qdm_gemm_w4ts.cu -o build/var_$name.o
objs=$(ls build/*.o | grep -v "var_\|qdm_gemm_w4ts.o\|qdm_gemm_trace.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libqdm_$name.so $objs build/var_$name.o
echo built libqdm_$name.so
