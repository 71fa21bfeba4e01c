#!/bin/bash
# build a kernel-variant library for A/B runs on the GPU box:  tools/build_variant.sh <name> "<-D flags>"
# -> gpurun_variants/lib_<name>.so  (load it with VOLPATH_B200_LIB=$PWD/gpurun_variants/lib_<name>.so)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; flags=$2
W=/tmp/vp_variant_$name
rm -rf $W; mkdir -p $W/pkg/csrc $ROOT/gpurun_variants
cp $ROOT/cuda-volpath_b200/csrc/*.cu $ROOT/cuda-volpath_b200/csrc/*.cuh $ROOT/cuda-volpath_b200/csrc/*.h $ROOT/cuda-volpath_b200/csrc/Makefile $W/pkg/csrc/
cp -r $ROOT/include $W/include
make -C $W/pkg/csrc -j8 EXTRA="$flags" OUT=$ROOT/gpurun_variants/lib_$name.so > $W/make.log 2>&1 || { tail -20 $W/make.log; exit 1; }
grep -A3 "k_render_fastILi2ELb0ELb1ELb0ELb0E" $W/pkg/csrc/volpath_render_fast.ptxas.log | grep -i "Used\|spill" | sed "s/^/[$name] /"
