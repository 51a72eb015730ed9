#!/bin/bash
# Builds two MPBoot binaries from the reference sources, out of tree, into integration/_bin/ (git-ignored):
#   mpboot-avx       the unmodified reference (upstream's own cmake line for the AVX target)
#   mpboot-avx-gpu   the same sources + integration/mpboot_gpu.patch + integration/mpgpu_shim.inc, linked against
#                    mpboot_b200/csrc/libmpgpu.so (RUNPATH $ORIGIN/../../mpboot_b200/csrc, so the pair travels together)
# The reference tree is copied to a scratch directory first (it is mounted read-only and the patch must not touch
# it); nothing of it is copied into this repository.  Usage: integration/build.sh [gpu|stock|all]   (default all)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(dirname "$HERE")"
REF="${MPBOOT_REF:-/root/reference}"
WORK="${MPBOOT_BUILD_DIR:-/tmp/mpboot-build}"
WHAT="${1:-all}"
JOBS="${JOBS:-8}"
[ -f "$REF/sprparsimony.cpp" ] || { echo "reference sources not found at $REF" >&2; exit 1; }
mkdir -p "$WORK" "$HERE/_bin"

copy_src() {   # $1 = destination
    rm -rf "$1"; mkdir -p "$1"
    (cd "$REF" && tar cf - --exclude=.git .) | (cd "$1" && tar xf -)
    chmod -R u+w "$1"
}
configure_and_make() {   # $1 = source dir, $2 = build dir, rest = extra cmake arguments
    local src="$1" bld="$2"; shift 2
    rm -rf "$bld"; mkdir -p "$bld"
    (cd "$bld" && cmake "$src" -DIQTREE_FLAGS=avx -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
        -DCMAKE_CXX_FLAGS="-std=gnu++11 -fpermissive -w" -DCMAKE_C_FLAGS="-w" "$@" > cmake.log 2>&1 \
        && make -j"$JOBS" > make.log 2>&1) || { tail -30 "$bld/make.log" >&2; exit 1; }
}

if [ "$WHAT" = all ] || [ "$WHAT" = stock ]; then
    copy_src "$WORK/src-stock"
    configure_and_make "$WORK/src-stock" "$WORK/build-stock"
    cp "$WORK/build-stock/mpboot-avx" "$HERE/_bin/mpboot-avx"
    echo "built $HERE/_bin/mpboot-avx"
fi
if [ "$WHAT" = all ] || [ "$WHAT" = gpu ]; then
    make -j"$JOBS" -C "$REPO/mpboot_b200/csrc" > /dev/null
    copy_src "$WORK/src-gpu"
    (cd "$WORK/src-gpu" && patch -p1 --binary < "$HERE/mpboot_gpu.patch")
    configure_and_make "$WORK/src-gpu" "$WORK/build-gpu" -DMPGPU_DIR="$REPO"
    cp "$WORK/build-gpu/mpboot-avx-gpu" "$HERE/_bin/mpboot-avx-gpu"
    echo "built $HERE/_bin/mpboot-avx-gpu"
fi
