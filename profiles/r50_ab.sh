#!/bin/bash
# A/B of contract50.cu builds (profiles/build_variants.sh): per-kernel ms at N=48, C=128 for each library given as argument
# usage: bash profiles/r50_ab.sh <tag> <batch> main pf2 pf4 ...
tag=$1; batch=$2; shift 2
for v in "$@"; do
  lib=profiles/_build/libccn_$v.so; [ "$v" = main ] && lib=graphflow_b200/libccn_b200.so
  CCN_B200_LIB=$PWD/$lib python profiles/r50_probe.py $batch > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err || tail -3 gpurun_out/${tag}_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_$v.json"))
print("$v", round(d["ms_per_step"],3), round(d["roofline_frac"],3), {k[4:]:round(x["ms_per_step"],3) for k,x in d["kernels"].items()})
PY
done
