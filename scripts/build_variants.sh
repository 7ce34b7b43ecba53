#!/bin/bash
# Build the experimental kernel variants (default-off macros) into sugar_b200/lib/variants/ for scripts/ab_bench.sh.
# Parity of a variant:  SGR_LIB_PATH=$PWD/sugar_b200/lib/variants/<lib>.so python -m pytest tests/test_gpu_parity.py -m gpu -q
#   scripts/build_variants.sh                 # the list below
#   scripts/build_variants.sh name "flags"    # one ad-hoc variant
set -e
cd "$(dirname "$0")/.."
mkdir -p sugar_b200/lib/variants
build() {  # name, flags
  SGR_LIB_OUT=sugar_b200/lib/variants/lib_$1.so SGR_NVCC_EXTRA="$2" python sugar_b200/build.py --force | tail -1
}
if [ $# -ge 1 ]; then
  build "$1" "$2"
else
  for v in "${VARIANTS[@]}"; do :; done
  while read -r name flags; do
    [ -z "$name" ] && continue
    build "$name" "$flags"
  done < scripts/variants.txt
fi
rm -f sugar_b200/lib/variants/*.o
ls -la sugar_b200/lib/variants
