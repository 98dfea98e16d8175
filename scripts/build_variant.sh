#!/bin/bash
# Build the working tree's libpnb200.so with extra nvcc flags for same-box A/B timing (dev tool):
#   scripts/build_variant.sh <tag> "<extra flags>"  ->  pyneapple_b200/csrc/_ab/libpnb200_<tag>.so  (PNB_LIB=...)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
mkdir -p "$T/pyneapple_b200" "$T/include"
cp -r "$ROOT/pyneapple_b200/csrc" "$T/pyneapple_b200/csrc"
cp "$ROOT/include/pyneapple_b200.h" "$T/include/"
rm -rf "$T/pyneapple_b200/csrc/_obj" "$T/pyneapple_b200/csrc/_ab" "$T/pyneapple_b200/csrc/libpnb200.so"
make -s -C "$T/pyneapple_b200/csrc" -j 16 "T1MODES=0" "EXTRA=$2" > "$T/build.log" 2>&1 || { tail "$T/build.log"; exit 1; }
mkdir -p "$ROOT/pyneapple_b200/csrc/_ab"
cp "$T/pyneapple_b200/csrc/libpnb200.so" "$ROOT/pyneapple_b200/csrc/_ab/libpnb200_$1.so"
rm -rf "$T"
echo "built _ab/libpnb200_$1.so with $2"
