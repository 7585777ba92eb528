#!/bin/bash
# oracle/ref_build.sh — TEST INFRASTRUCTURE ONLY.
# Builds oracle/_ref/libref.so: the reference's own hot-path sources, compiled UNMODIFIED from where they lie under
# /root/reference (nothing is copied into this repository), against the OpenCV / Eigen stand-ins in oracle/refshim/
# (OpenCV 3.4 C++, opencv_contrib, Eigen, g2o and Pangolin are not installed in this image and there is no network, so the
# reference's CMake build cannot run; these translation units need only the API slice the stand-ins provide).
# Linked with oracle/ref_harness.cpp (C entry points) and liboracle.so (the cv2-pinned image primitives the stand-in forwards to).
# Output only under oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).  Usage: bash oracle/ref_build.sh [-f]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${SSLPL_REFERENCE_DIR:-/root/reference}
OUT="$HERE/_ref"
mkdir -p "$OUT/obj"
if [ ! -d "$REF/src" ]; then echo "ref_build: $REF not present (GPU box) - using the prebuilt $OUT/libref.so" >&2; exit 0; fi
make -C "$HERE" -s
# the system compiler (dynamic libstdc++.so.6); $CXX of this image points at a toolchain that links libstdc++ statically, whose stream
# number formatting crashes inside a dlopen()ed library
CXX=${SSLPL_REF_CXX:-/usr/bin/g++}
# -ffp-contract=off: the canonical no-FMA definition, as for the oracle itself
FLAGS="-std=c++14 -O2 -ffp-contract=off -fno-fast-math -march=x86-64-v3 -fPIC -w -I$HERE/refshim -I$REF/include -I$REF -I$HERE"
SRCS="src/ORBextractor.cc src/ORBmatcher.cc src/LSDmatcher.cpp src/ExtractLineSegment.cpp src/Frame.cc src/KeyFrame.cc src/MapPoint.cc
      src/MapLine.cpp src/Map.cc src/KeyFrameDatabase.cc Thirdparty/DBoW2/DBoW2/FORB.cpp Thirdparty/DBoW2/DBoW2/BowVector.cpp
      Thirdparty/DBoW2/DBoW2/FeatureVector.cpp Thirdparty/DBoW2/DBoW2/ScoringObject.cpp Thirdparty/DBoW2/DUtils/Random.cpp
      Thirdparty/DBoW2/DUtils/Timestamp.cpp"
newer() { [ ! -e "$2" ] || [ "$1" -nt "$2" ]; }
SHIMSTAMP=$(find "$HERE/refshim" "$HERE/oracle.h" -type f -newer "$OUT/libref.so" 2>/dev/null | head -1)
OBJS=""
pids=""
for s in $SRCS; do
  o="$OUT/obj/$(echo $s | tr '/' '_').o"
  OBJS="$OBJS $o"
  if [ "$1" = "-f" ] || [ -n "$SHIMSTAMP" ] || newer "$REF/$s" "$o"; then ( $CXX $FLAGS -c "$REF/$s" -o "$o" ) & pids="$pids $!"; fi
done
for s in refshim/minicv.cpp ref_harness.cpp; do
  o="$OUT/obj/$(echo $s | tr '/' '_').o"
  OBJS="$OBJS $o"
  if [ "$1" = "-f" ] || [ -n "$SHIMSTAMP" ] || newer "$HERE/$s" "$o"; then ( $CXX $FLAGS -c "$HERE/$s" -o "$o" ) & pids="$pids $!"; fi
done
for p in $pids; do wait $p; done
# -Bsymbolic-functions: operator new/delete of ref_harness.cpp (malloc, or the bump arena inside a BumpScope) serve this library only
$CXX -shared -o "$OUT/libref.so" $OBJS -Wl,-Bsymbolic-functions -Wl,--no-undefined \
     -L"$HERE" -loracle -Wl,-rpath,'$ORIGIN/..' -lpthread -lm
echo "ref_build: $OUT/libref.so ($(echo $SRCS | wc -w) reference translation units)"
