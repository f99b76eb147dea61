#!/bin/bash
# build one library variant of the kernel: tools/build_variant.sh NAME -DSEDI_...=v ...   -> build_variants/NAME.so (+ NAME.ptxas with registers / spills)
name=$1; shift
mkdir -p build_variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v "$@" \
  -o build_variants/$name.so sedifoam_b200/csrc/sedi_engine.cu 2> build_variants/$name.ptxas
rc=$?
grep -A1 "k_step_sellILi3ELb1ELi0" build_variants/$name.ptxas | grep -E "Used|spill" | tr '\n' ' ' | sed 's/ptxas info    ://g'; echo " [$name rc=$rc]"
grep -E "error" build_variants/$name.ptxas | head -5
exit $rc
