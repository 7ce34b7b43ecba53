#!/bin/bash
# Build the experimental kernel variants (default-off macros) into sugar_b200/lib/variants/ for scripts/ab_bench.sh.
# Parity of a variant:  SGR_LIB_PATH=$PWD/sugar_b200/lib/variants/<lib>.so python -m pytest tests/test_gpu_parity.py -m gpu -q
set -e
cd "$(dirname "$0")/.."
mkdir -p sugar_b200/lib/variants
build() {  # name, flags
  SGR_LIB_OUT=sugar_b200/lib/variants/lib_$1.so SGR_NVCC_EXTRA="$2" python sugar_b200/build.py --force | tail -1
}
build bwd_one_barrier            "-DSGR_BWD_ONE_BARRIER=1"
build bwd_balanced               "-DSGR_BWD_BALANCED_LOADERS=1"
build bwd_one_barrier_balanced   "-DSGR_BWD_ONE_BARRIER=1 -DSGR_BWD_BALANCED_LOADERS=1"
build bwd_butterfly              "-DSGR_BWD_CHUNKED=0"
build fwd_pipelined              "-DSGR_FWD_PIPELINED=1"
rm -f sugar_b200/lib/variants/*.o
ls -la sugar_b200/lib/variants
