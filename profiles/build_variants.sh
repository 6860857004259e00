#!/bin/bash
# A/B builds of the library with pieces of contract18_fused.cu compiled out (profiles/_build/libccn_*.so, selected with CCN_B200_LIB).
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/_build
python -m graphflow_b200.build >/dev/null
OBJ=graphflow_b200/csrc/_obj
for v in ${VARIANTS:-"stage:-DCCN_STAGE_ROLLED" "list:-DCCN_LIST_NOINLINE" "both:-DCCN_STAGE_ROLLED -DCCN_LIST_NOINLINE"}; do
  name=${v%%:*}; flags=${v#*:}
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags \
      -c -o profiles/_build/fused_$name.o graphflow_b200/csrc/contract18_fused.cu
  objs=$(ls $OBJ/*.o | grep -v contract18_fused.o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o profiles/_build/libccn_$name.so $objs profiles/_build/fused_$name.o
done
ls -la profiles/_build/*.so
