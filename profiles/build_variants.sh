#!/bin/bash
# A/B builds of the library (profiles/_build/libccn_<name>.so, selected with CCN_B200_LIB): VARIANTS='name:-DFLAG@-DFLAG2 ...' rebuilds
# SRC (default contract18_fused_bwd.cu) with the flags and links it with the other objects of the regular build.
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/_build
python -m graphflow_b200.build >/dev/null
OBJ=graphflow_b200/csrc/_obj
for v in ${VARIANTS:-"stage:-DCCN_STAGE_ROLLED" "list:-DCCN_LIST_NOINLINE" "both:-DCCN_STAGE_ROLLED -DCCN_LIST_NOINLINE"}; do
  name=${v%%:*}; flags=$(echo ${v#*:} | tr @ " ")
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags \
      -c -o profiles/_build/fused_$name.o graphflow_b200/csrc/${SRC:-contract18_fused_bwd.cu}
  objs=$(ls $OBJ/*.o | grep -v $(basename ${SRC:-contract18_fused_bwd.cu} .cu).o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o profiles/_build/libccn_$name.so $objs profiles/_build/fused_$name.o
done
ls -la profiles/_build/*.so
