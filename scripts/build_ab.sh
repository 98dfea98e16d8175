#!/bin/bash
# Build libpnb200.so of another commit next to the working tree's one for same-box A/B timing (dev tool):
#   scripts/build_ab.sh <commit> <tag>   ->  pyneapple_b200/csrc/_ab/libpnb200_<tag>.so   (use with PNB_LIB=...)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
git -C "$ROOT" archive "$1" pyneapple_b200/csrc include | tar -x -C "$T"
make -s -C "$T/pyneapple_b200/csrc" -j 16 "T1MODES=0" > "$T/build.log" 2>&1 || { tail "$T/build.log"; exit 1; }
mkdir -p "$ROOT/pyneapple_b200/csrc/_ab"
cp "$T/pyneapple_b200/csrc/libpnb200.so" "$ROOT/pyneapple_b200/csrc/_ab/libpnb200_$2.so"
rm -rf "$T"
echo "built _ab/libpnb200_$2.so from $1"
