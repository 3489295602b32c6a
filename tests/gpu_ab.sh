#!/bin/bash
# A/B of a compile-time switch ON THE GPU BOX: gpurun -- 'bash tests/gpu_ab.sh -DBK_NO_PREFETCH C5'
cd "$(dirname "$0")/.."
FLAG=$1; CFG=${2:-C2}
for f in "" "$FLAG"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off $f \
    -Iinclude -Ipymc_bart_b200/csrc -shared -o pymc_bart_b200/libpgbart_b200.so pymc_bart_b200/csrc/pgbart_b200.cu pymc_bart_b200/csrc/pgbart_predict.cu 2>/dev/null
  echo "flags: [$f]"; python bench.py --steps 100 --warmup 10 --profile-only --config $CFG; python bench.py --steps 100 --warmup 10 --profile-only --config $CFG
done
