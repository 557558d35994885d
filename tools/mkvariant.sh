#!/bin/bash
# Build the current tree as variants/lib_<name>.so (optionally with extra nvcc flags, e.g. EXTRA=-DFT_MINCTAS=5) for tools/ab.sh.
#     tools/mkvariant.sh <name> [file.cu=replacement.cu ...]
N=$1; shift
T=$(mktemp -d)
mkdir -p $T/a/b $T/a/include
cp -r object_slam_b200/csrc $T/a/b/csrc; rm -rf $T/a/b/csrc/build
cp include/obslam_b200.h $T/a/include/
for kv in "$@"; do cp "${kv#*=}" "$T/a/b/csrc/${kv%%=*}"; done
make -s -j 16 -C $T/a/b/csrc EXTRA="$EXTRA" > $T/log 2>&1 || { tail -5 $T/log; echo build failed; exit 1; }
mkdir -p variants; cp $T/a/b/libobslam_b200.so variants/lib_$N.so; rm -rf $T; ls -la variants/lib_$N.so
