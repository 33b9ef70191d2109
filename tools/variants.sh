#!/bin/bash
# Tuning helper: build variant libraries of the same sources with -D flags (here, no GPU needed) and bench each on the GPU box.
#   tools/variants.sh build  NAME "-DFLAG=1 ..." [NAME2 "..."]...     -> build/var_NAME.so
#   tools/variants.sh bench  [extra bench.py args]                      -> gpurun_out/variants.txt (run under gpurun)
set -e
cd "$(dirname "$0")/.."
CS=decombinator_b200/csrc
if [ "$1" == "build" ]; then
  shift; mkdir -p build
  while [ $# -ge 2 ]; do
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3 -shared $2 \
      -I include -o build/var_$1.so $CS/decombine.cu $CS/collapse.cu $CS/tagset.cpp $CS/pack.cpp $CS/fastq.cpp $CS/synth.cpp $CS/error.cpp -lpthread &
    shift 2
  done
  wait; ls -la build/
else
  shift || true
  mkdir -p gpurun_out; : > gpurun_out/variants.txt
  for f in build/var_*.so; do
    for rep in 1 2; do
      DCB_LIB=$PWD/$f timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$f', '%.2f G reads/s  exact %.4f ms  general %.3f ms  e2e %.3f G' % (d['value']/1e9, d['kernels_ms']['dcb_exact_kernel'], d['kernels_ms']['dcb_general_kernel'], d['e2e']['value']/1e9))" >> gpurun_out/variants.txt || echo "$f FAILED" >> gpurun_out/variants.txt
    done
  done
  cat gpurun_out/variants.txt
fi
