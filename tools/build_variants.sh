#!/bin/bash
# Build experiment variants of the library: tools/build_variants.sh name "-DFLAG ..." [name "-D..."]...
# -> sucem-fem_b200/lib/variants/lib_<name>.so (git-ignored, travels to the GPU box), used with SUCEM_B200_LIB=...
set -e
cd "$(dirname "$0")/../sucem-fem_b200"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  make -s -j8 BUILD=build_$name LIB=lib/variants/lib_$name.so EXTRA="$flags" 2>&1 | grep -E "error|spill" || true
  cuobjdump -res-usage build_$name/assemble_canonical.o 2>/dev/null | grep -A1 "k_canon_edges" | grep -o "REG:[0-9]* STACK:[0-9]*" | sed "s/^/$name edges /"
done
