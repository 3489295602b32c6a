#!/bin/bash
# A/B of library variants built beforehand (build_variants/lib_<name>.so, nvcc on the dev box) ON THE GPU BOX:
#   gpurun -- 'bash tests/gpu_ab_prebuilt.sh base wacq both'
# prints the device-timed C2 (twice) and C5 numbers per variant; the snapshot on the box is scratch.
cd "$(dirname "$0")/.."
cp pymc_bart_b200/libpgbart_b200.so /tmp/lib_keep.so
for n in "$@"; do
  cp build_variants/lib_$n.so pymc_bart_b200/libpgbart_b200.so
  echo "=== variant: $n"
  for r in 1 2; do python bench.py --steps 128 --warmup 10 --profile-only --config C2 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C2', round(d['value'],1), {k: round(v) for k, v in d['in_kernel_us'].items() if v is not None})"; done
  python bench.py --steps 32 --warmup 5 --profile-only --config C5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C5', round(d['value'],2), {k: round(v) for k, v in d['in_kernel_us'].items() if v is not None})"
done
cp /tmp/lib_keep.so pymc_bart_b200/libpgbart_b200.so
