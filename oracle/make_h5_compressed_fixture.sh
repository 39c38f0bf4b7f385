#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds GATB-core from /root/reference in a scratch tree (like build_ref.sh), compiles
# oracle/h5_compressed_main.cpp against it and writes tests/golden/cfg1_small_compressed.h5.gz: the graph file of the
# cfg1_small case with shuffle + deflate on every dataset (GATB's -out-compress).  Runs only where /root/reference is.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${G2S_REFERENCE:-/root/reference}"
BLD="$HERE/_ref/build_h5"
mkdir -p "$BLD"
if [ ! -f "$BLD/gatb/lib/Release/libgatbcore.a" ]; then
  rm -rf "$BLD/gatb-core-src" "$BLD/gatb"
  cp -r "$REF/thirdparty/gatb-core/gatb-core" "$BLD/gatb-core-src"
  chmod -R u+w "$BLD/gatb-core-src"
  mkdir -p "$BLD/gatb"
  (cd "$BLD/gatb" && cmake -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_BUILD_TYPE=Release \
      -DGATB_CORE_EXCLUDE_TOOLS=1 -DGATB_CORE_EXCLUDE_TESTS=1 -DGATB_CORE_EXCLUDE_EXAMPLES=1 \
      "$BLD/gatb-core-src" > cmake.log 2>&1 && make -j"${JOBS:-8}" > make.log 2>&1)
fi
INC="-I$BLD/gatb-core-src/src -I$BLD/gatb/include -I$BLD/gatb/include/Release -I$BLD/gatb-core-src/thirdparty -I$BLD/gatb/include/Release/hdf5"
LIB="-L$BLD/gatb/lib/Release -lgatbcore -lhdf5 -ldl -lz -lpthread"
g++ -std=c++11 -O2 -w $INC "$HERE/h5_compressed_main.cpp" $LIB -o "$BLD/h5_compressed"
WORK="$BLD/work"; rm -rf "$WORK"; mkdir -p "$WORK"
python - "$WORK" <<'P'
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(sys.argv[0])) if False else os.getcwd(), "tests"))
sys.path.insert(0, os.getcwd())
import cases
w = cases.build(cases.CASE_BY_NAME["cfg1_small"])
w.write(sys.argv[1])
P
(cd "$WORK" && "$BLD/h5_compressed" reads.fa 31 2 6 "$WORK/graph" > run.log 2>&1)
python - "$WORK/graph.h5" "$HERE/../tests/golden/cfg1_small_compressed.h5.gz" <<'P'
import gzip, sys
raw = open(sys.argv[1], "rb").read()
open(sys.argv[2], "wb").write(gzip.compress(raw, 9, mtime=0))
print(f"compressed graph file: {len(raw)} bytes -> {len(gzip.compress(raw, 9, mtime=0))} gzip'ed")
P
rm -rf "$BLD"
