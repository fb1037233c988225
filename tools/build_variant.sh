#!/bin/bash
# Build an alternative libpea_b200 (same ABI) with extra -D flags on ONE source file, for A/B timing runs:
#   tools/build_variant.sh <name> <source.cu> [extra nvcc flags...]   ->  build_ab/lib<name>.so
# The other objects are taken from the regular in-tree build (run `make -C pixel-embedded-affinity_b200/csrc` first).
# Use with:  PEA_B200_LIB=build_ab/lib<name>.so python tools/k4_ab.py cfg3
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
CSRC="$ROOT/pixel-embedded-affinity_b200/csrc"
name="$1"; src="$2"; shift 2
mkdir -p "$ROOT/build_ab"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
obj="$ROOT/build_ab/${name}_$(basename "$src" .cu).o"
$NVCC $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I"$ROOT/include" -I"$CSRC" "$@" -c "$CSRC/$src" -o "$obj"
others=$(ls "$CSRC"/pea_*.o | grep -v "/$(basename "$src" .cu).o")
$NVCC $ARCH -shared -o "$ROOT/build_ab/lib${name}.so" "$obj" $others
echo "built build_ab/lib${name}.so"
