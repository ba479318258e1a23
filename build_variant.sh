#!/bin/bash
# A/B builds of libmavi_cuda.so with extra nvcc flags: ./build_variant.sh <name> <flags...>  ->  build/var_<name>/libmavi_cuda.so
# (run with MAVI_LIB_PATH=build/var_<name>/libmavi_cuda.so; build/ is git-ignored but travels with gpurun)
set -e
name=$1; shift
src=$(dirname "$0")/mavi.jl_b200/csrc
tmp=/tmp/mavi_var_$name
rm -rf $tmp; mkdir -p $tmp/mavi.jl_b200 $tmp/include
cp -r $src $tmp/mavi.jl_b200/csrc; cp $(dirname "$0")/include/*.h $tmp/include/
rm -f $tmp/mavi.jl_b200/csrc/*.o $tmp/mavi.jl_b200/csrc/*.so
make -C $tmp/mavi.jl_b200/csrc -j6 EXTRA="$*" > $tmp/build.log 2>&1 || { tail -20 $tmp/build.log; exit 1; }
mkdir -p $(dirname "$0")/build/var_$name
cp $tmp/mavi.jl_b200/csrc/libmavi_cuda.so $(dirname "$0")/build/var_$name/
echo built build/var_$name/libmavi_cuda.so
