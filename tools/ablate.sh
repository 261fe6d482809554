#!/bin/bash
# Builds debug copies of the library with one part of k_project left out each (wrong results, right timing of
# the rest) and times a3d_project with each on one workload: what every part of the kernel costs.
#   bash tools/ablate.sh build      (in the build container)
#   bash tools/ablate.sh run c3_mini (on the GPU box)
cd "$(dirname "$0")/.."
FLAGS="-shared -Xcompiler -fPIC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -ffp-contract=off -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -I include"
if [ "$1" = build ]; then
  mkdir -p tools/_build
  for v in RED EXACT_LIST PHASE_A WRITE; do
    nvcc $FLAGS -DA3D_ABLATE_$v -o tools/_build/liba3d_ablate_$v.so articulation3d_b200/csrc/a3d.cu || exit 1
  done
  nvcc $FLAGS -DA3D_ABLATE_PHASE_A -DA3D_ABLATE_EXACT_LIST -DA3D_ABLATE_WRITE -o tools/_build/liba3d_ablate_ALL.so articulation3d_b200/csrc/a3d.cu || exit 1
  ls -la tools/_build/*.so
else
  wl=${2:-c3_mini}
  echo "== shipped"; AB_ITERS=10 AB_OUT_MODE=1 python tools/project_ab.py $wl 2>&1 | grep filter | tail -1
  for v in RED EXACT_LIST PHASE_A WRITE ALL; do
    echo "== without $v"; A3D_LIB=$PWD/tools/_build/liba3d_ablate_$v.so AB_ITERS=10 AB_OUT_MODE=1 python tools/project_ab.py $wl 2>&1 | grep filter | tail -1
  done
fi
