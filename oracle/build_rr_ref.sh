#!/bin/bash
# oracle/_ref/rr: the reference's own read_recruitment filter (rr.cpp + its vendored edlib + kseq), compiled from the
# sources where they lie under /root/reference with the reference's own flags (scripts/read_recruitment/Makefile:2-4);
# outputs only into oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot).  TEST INFRASTRUCTURE.
set -e
REF=/root/reference/scripts/read_recruitment
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
[ -d "$REF" ] || { echo "build_rr_ref.sh: $REF is not here; keeping whatever $OUT holds"; exit 0; }
mkdir -p "$OUT"
g++ --std=c++14 -O2 -c "$REF/edlib/src/edlib.cpp" -o "$OUT/edlib.o" -I "$REF/edlib/include"
g++ --std=c++14 -O2 -c "$REF/rr.cpp" -o "$OUT/rr.o" -I "$REF/edlib/include" -I "$REF"
g++ "$OUT/rr.o" "$OUT/edlib.o" -o "$OUT/rr" -lz
rm -f "$OUT/rr.o" "$OUT/edlib.o"
echo "oracle/_ref/rr built"
