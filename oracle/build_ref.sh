#!/usr/bin/env bash
# Build the UNMODIFIED reference (rikuu/Gap2Seq @ /root/reference) into
# oracle/_ref/ so tests and bench.py can use it as the parity oracle and the
# CPU baseline.  TEST INFRASTRUCTURE ONLY - the product never links or runs it.
#
#   oracle/_ref/bin/Gap2Seq-core        reference binary   (-O2)
#   oracle/_ref/bin/Gap2Seq-core-stats  same with -DSTATS  (prints DP cell counts)
#   oracle/_ref/bin/GapCutter, GapMerger
#   oracle/_ref/bin/Gap2Seq-core-g2s    INTEGRATION.md form 2: the reference's Gap2Seq.cpp + GATB's graph with the two
#                                       fill_gap call sites routed to libg2s_gpu.so through integration/g2s_gatb_glue.cpp
#                                       (needs a GPU to run; built only when gap2seq_b200/lib/libg2s_gpu.so exists)
#
# Sources are compiled where they lie under /root/reference; only GATB-core
# (vendored third-party library, needs its own CMake because of the bundled
# HDF5) is copied to a scratch build tree under oracle/_ref/build/ (listed in
# .gpurunignore, it does not travel).  Boost.Graph is absent from the image:
# oracle/boost_shim/ supplies the handful of calls Gap2Seq.cpp makes.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${G2S_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
BLD="$OUT/build"
JOBS="${JOBS:-8}"

if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present - keeping prebuilt oracle/_ref as is" >&2
  exit 0
fi
G2SLIB="$HERE/../gap2seq_b200/lib"
if [ -x "$OUT/bin/Gap2Seq-core" ] && [ -x "$OUT/bin/Gap2Seq-core-stats" ] && [ "${FORCE:-0}" != 1 ] && \
   { [ -x "$OUT/bin/Gap2Seq-core-g2s" ] || [ ! -f "$G2SLIB/libg2s_gpu.so" ]; }; then
  echo "build_ref: oracle/_ref/bin already built" >&2
  exit 0
fi

mkdir -p "$BLD" "$OUT/bin"
if [ ! -f "$BLD/gatb/lib/Release/libgatbcore.a" ]; then
  rm -rf "$BLD/gatb-core-src" "$BLD/gatb"
  cp -r "$REF/thirdparty/gatb-core/gatb-core" "$BLD/gatb-core-src"
  chmod -R u+w "$BLD/gatb-core-src"
  mkdir -p "$BLD/gatb"
  (cd "$BLD/gatb" && cmake -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_BUILD_TYPE=Release \
      -DGATB_CORE_EXCLUDE_TOOLS=1 -DGATB_CORE_EXCLUDE_TESTS=1 -DGATB_CORE_EXCLUDE_EXAMPLES=1 \
      "$BLD/gatb-core-src" > cmake.log 2>&1 && make -j"$JOBS" > make.log 2>&1)
fi

INC="-I$HERE/boost_shim -I$BLD/gatb-core-src/src -I$BLD/gatb/include -I$BLD/gatb/include/Release -I$BLD/gatb-core-src/thirdparty -I$BLD/gatb/include/Release/hdf5"
LIB="-L$BLD/gatb/lib/Release -lgatbcore -lhdf5 -ldl -lz -lpthread"
CXXF="-std=c++11 -O2 -w"
g++ $CXXF $INC "$REF/src/Gap2Seq.cpp" "$REF/src/main.cpp" $LIB -o "$OUT/bin/Gap2Seq-core"
g++ $CXXF -DSTATS $INC "$REF/src/Gap2Seq.cpp" "$REF/src/main.cpp" $LIB -o "$OUT/bin/Gap2Seq-core-stats"
g++ $CXXF $INC "$REF/src/GapCutter.cpp" $LIB -o "$OUT/bin/GapCutter"
g++ $CXXF $INC "$REF/src/GapMerger.cpp" $LIB -o "$OUT/bin/GapMerger"
# form 2: the reference source is patched in a scratch copy (never committed): one #include and the two call sites
if [ -f "$G2SLIB/libg2s_gpu.so" ]; then
  mkdir -p "$BLD/form2"
  sed -e 's/= fill_gap(graph,/= g2s_fill_gap_graph(graph,/' \
      -e 's/#include "Gap2Seq.hpp"/#include "Gap2Seq.hpp"\n#include "g2s_gatb_glue.h"/' "$REF/src/Gap2Seq.cpp" > "$BLD/form2/Gap2Seq.cpp"
  test "$(grep -c 'g2s_fill_gap_graph(graph,' "$BLD/form2/Gap2Seq.cpp")" = 2
  g++ $CXXF $INC -I"$REF/src" -I"$HERE/../integration" -I"$HERE/../include" "$BLD/form2/Gap2Seq.cpp" "$REF/src/main.cpp" \
      "$HERE/../integration/g2s_gatb_glue.cpp" $LIB -L"$G2SLIB" -lg2s_gpu -Wl,-rpath,'$ORIGIN/../../../gap2seq_b200/lib' \
      -o "$OUT/bin/Gap2Seq-core-g2s"
fi
# the GATB build tree (200 MB of objects) is scratch: only the binaries stay
rm -rf "$BLD"
echo "build_ref: done -> $OUT/bin" >&2
