#!/bin/bash
# Builds the UNMODIFIED reference (steineggerlab/Metabuli, /root/reference) into oracle/_ref/metabuli.
# Container-only (needs /root/reference and cmake); the binary is git-ignored but travels to the GPU box with the snapshot,
# where bench.py uses it as the timed CPU baseline (`cpu_baseline.kind = "reference"`) and as the checker of the
# bit-exact parity sample, and tests/golden/gen_golden.sh uses it to write the golden vectors.  Sources are compiled
# from a scratch copy under /tmp (the reference tree is read-only and its cmake build writes generated files); nothing
# but the resulting binary enters the repo.  Recipe: SURVEY.md §8c.
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ -x "$OUT/metabuli" ] && [ -x "$OUT/ref_mask" ] && [ "${1:-}" != "force" ]; then echo "oracle/_ref/metabuli and ref_mask already built"; exit 0; fi
if [ ! -d /root/reference ]; then echo "no /root/reference here: cannot build oracle/_ref"; exit 0; fi
W=${MBL_REF_BUILD_DIR:-/tmp/oracle}
if [ ! -x $W/build/src/metabuli ]; then
  rm -rf $W && mkdir -p $W && cp -r /root/reference $W/src && chmod -R u+w $W/src && mkdir -p $W/build && cd $W/build
  # without the two compiler flags cmake picks /opt/gcc/bin/g++, whose wrapper cannot find libgomp.spec
  cmake -DCMAKE_BUILD_TYPE=Release -DCMAKE_C_COMPILER=/usr/bin/gcc -DCMAKE_CXX_COMPILER=/usr/bin/g++ ../src > $W/cmake.log 2>&1
  make -j"$(nproc)" metabuli > $W/make.log 2>&1
fi
mkdir -p "$OUT"
cp $W/build/src/metabuli "$OUT/metabuli"
strip "$OUT/metabuli" || true
# oracle/ref_mask_main.cpp: the reference's own NucleotideMatrix / ProbabilityMatrix / tantan objects behind a line filter
# (golden generator of the --mask 1 cases); compiled with the flags and include paths of the reference's own build
FL=$W/build/src/CMakeFiles/metabuli.dir/flags.make
INC=$(grep '^CXX_INCLUDES' $FL | sed 's/CXX_INCLUDES = //'); DEF=$(grep '^CXX_DEFINES' $FL | sed 's/CXX_DEFINES = //')
L=$W/build/lib/mmseqs
/usr/bin/g++ -O3 -DNDEBUG -fsigned-char -march=native -std=c++1y -fopenmp $DEF $INC -c "$HERE/ref_mask_main.cpp" -o $W/ref_mask.o
/usr/bin/g++ -fopenmp $W/ref_mask.o -o "$OUT/ref_mask" $L/src/libmmseqs-framework.a $W/build/src/version/libversion.a -latomic \
  $L/lib/tinyexpr/libtinyexpr.a $L/lib/zstd/build/cmake/lib/libzstd.a $L/lib/microtar/libmicrotar.a $L/lib/tantan/libtantan.a -lz -lpthread
strip "$OUT/ref_mask" || true
ls -la "$OUT/metabuli" "$OUT/ref_mask"
