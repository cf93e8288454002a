#!/bin/bash
# build experimental variants in parallel: bash tools/build_variants.sh name1="-DX=1 -DY=2" name2="..."
mkdir -p build/var
pids=()
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  ( MBE_NVCC_EXTRA="$flags" MBE_LIB_OUT=$PWD/build/var/$name.so python mbelib-neo_b200/build.py --force > build/var/$name.log 2>&1 || echo "BUILD FAILED $name" ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
ls -la build/var/*.so
