#!/bin/bash
# A/B of compile-time variants ON THE GPU BOX (the snapshot there is scratch):
#   gpurun -- 'bash tests/gpu_variants.sh "-DBK_CTA_THREADS=1024" "-DBK_CTA_THREADS=768" "-DBK_CTA_THREADS=512"'
# Each argument is one set of nvcc flags; prints the device-timed C2 and C5 numbers (bench.py --profile-only) per variant,
# then restores the default build.
cd "$(dirname "$0")/.."
build() {
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off $1 \
    -Iinclude -Ipymc_bart_b200/csrc -shared -o pymc_bart_b200/libpgbart_b200.so pymc_bart_b200/csrc/pgbart_b200.cu pymc_bart_b200/csrc/pgbart_predict.cu 2>/dev/null
}
for f in "$@"; do
  build "$f"
  echo "=== flags: [$f]"
  python bench.py --steps 100 --warmup 10 --profile-only --config C2
  python bench.py --steps 100 --warmup 10 --profile-only --config C2
  python bench.py --steps 30 --warmup 5 --profile-only --config C5
  [ -n "$AB_MORE" ] && python bench.py --steps 50 --warmup 5 --profile-only --config C3 && python bench.py --steps 50 --warmup 5 --profile-only --config C4
done
build ""
