#!/bin/bash
# Build the library of another git revision (or of the working tree with extra nvcc flags) into build_variants/
# for a same-box A/B run with tools/ab_probe.sh (RMB_LIB):
#   tools/ab_build.sh <git-rev> [name]        -> build_variants/lib_<name or rev>.so
#   EXTRA="-DRMB_VEC_MINB=4" tools/ab_build.sh WORKTREE minb4
# build_variants/ is git-ignored but travels to the GPU box with the gpurun snapshot.
set -e
rev=$1
name=${2:-$rev}
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/build_variants"
flags="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared $EXTRA"
if [ "$rev" = WORKTREE ]; then
  src="$root/richmol_b200/csrc"
else
  tmp=$(mktemp -d)
  git -C "$root" archive "$rev" richmol_b200/csrc include | tar -x -C "$tmp"
  src="$tmp/richmol_b200/csrc"
fi
(cd "$src" && nvcc $flags -o "$root/build_variants/lib_$name.so" rmb.cu -lcudart)
[ "$rev" = WORKTREE ] || rm -rf "$tmp"
ls -la "$root/build_variants/lib_$name.so"
