#!/bin/bash
# Builds the -DBK_PROFILE_CTRL variant of the library ON THE GPU BOX (the snapshot there is scratch) and prints the
# in-kernel phase timers:  gpurun -- 'bash tests/gpu_prof.sh C2 40 1; bash tests/gpu_prof.sh C2 40 4'
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -DBK_PROFILE_CTRL \
  -Iinclude -Ipymc_bart_b200/csrc -shared -o pymc_bart_b200/libpgbart_b200.so pymc_bart_b200/csrc/pgbart_b200.cu pymc_bart_b200/csrc/pgbart_predict.cu 2>/dev/null
python tests/gpu_profile_phases.py "$@"
