#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY. Builds the UNMODIFIED reference (alevar/tiebrush) from the sources
# where they lie under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box).
# Outputs: oracle/_ref/{tiebrush,tiecov,htsfile} (-O2, the README's Release intent), tiebrush_O0 / tiecov_O0 (-O0 -g: what the
# reference's CMakeLists.txt:49 actually produces), tiewrap.py (the reference's own parallel driver, placed next to the binary
# as its install rule does), hts_tool (oracle/hts_tool.c: SAM->BAM converter and decode-only pass over the vendored htslib).  Nothing from the reference is copied into the repo;
# the build happens on a scratch copy under /tmp because the reference tree is read-only and
# htslib's Makefile writes objects next to its sources.
# Recipe follows SURVEY.md §8c (the reference's own CMake needs network + autoreconf, so it is not used).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present (GPU box?) - using prebuilt $OUT if any" >&2
  exit 0
fi
if [ -x "$OUT/tiebrush" ] && [ -x "$OUT/tiecov" ] && [ -x "$OUT/htsfile" ] && [ -x "$OUT/tiebrush_O0" ] && [ -x "$OUT/hts_tool" ] && [ -f "$OUT/tiewrap.py" ] \
   && [ "$OUT/hts_tool" -nt "$HERE/hts_tool.c" ] && [ "${FORCE:-0}" != 1 ]; then
  exit 0
fi
W=$(mktemp -d /tmp/refbuild.XXXXXX)
trap 'rm -rf "$W"' EXIT
cp -r "$REF/src" "$REF/include" "$W/"
chmod -R u+w "$W"
cd "$W/include/htslib"
# no bz2/lzma/curl headers in this image: minimal config.h, CRAM codecs that need them are compiled out
printf '#ifndef _XOPEN_SOURCE\n#define _XOPEN_SOURCE 600\n#endif\n#define HAVE_DRAND48 1\n' > config.h
echo '#define HTSCODECS_VERSION_TEXT "1.5.0"' > htscodecs/htscodecs/version.h
make -j"$(nproc)" lib-static NONCONFIGURE_OBJS= >/dev/null 2>&1
gcc -O2 -I. -o htsfile htsfile.c libhts.a -lz -lpthread -lm
cd "$W"
SRC="src/GSam.cpp src/tmerge.cpp include/gclib/GStr.cpp include/gclib/GArgs.cpp include/gclib/GBase.cpp"
CXXF="-std=c++11 -fpermissive -w -DNOCURL=1 -O2 -Iinclude -Iinclude/htslib"
g++ $CXXF src/tiebrush.cpp $SRC include/htslib/libhts.a -lz -lpthread -o tiebrush
CXXF0="-std=c++11 -fpermissive -w -DNOCURL=1 -O0 -g -Iinclude -Iinclude/htslib"
g++ $CXXF0 src/tiebrush.cpp $SRC include/htslib/libhts.a -lz -lpthread -o tiebrush_O0
gcc -O2 -Iinclude/htslib -Iinclude "$HERE/hts_tool.c" include/htslib/libhts.a -lz -lpthread -lm -o hts_tool
# tiecov includes <libBigWig/bigWig.h>, which the reference fetches at CMake time (not vendored):
# a declaration-only stub satisfies the compiler; -W (BigWig) is out of scope and never exercised.
mkdir -p stub/libBigWig
cat > stub/libBigWig/bigWig.h <<'EOS'
#pragma once
#include <stdint.h>
typedef struct { void* cl; } bigWigFile_t;
static inline int bwInit(size_t){return 1;}
static inline bigWigFile_t* bwOpen(char*, void*, const char*){return 0;}
static inline int bwCreateHdr(bigWigFile_t*, int32_t){return 1;}
static inline void* bwCreateChromList(char**, uint32_t*, int64_t){return 0;}
static inline int bwWriteHdr(bigWigFile_t*){return 1;}
static inline int bwAddIntervals(bigWigFile_t*, char**, uint32_t*, uint32_t*, float*, uint32_t){return 1;}
static inline int bwAppendIntervals(bigWigFile_t*, uint32_t*, uint32_t*, float*, uint32_t){return 1;}
static inline void bwClose(bigWigFile_t*){}
static inline void bwCleanup(void){}
EOS
g++ $CXXF -Istub src/tiecov.cpp $SRC include/htslib/libhts.a -lz -lpthread -o tiecov
g++ $CXXF0 -Istub src/tiecov.cpp $SRC include/htslib/libhts.a -lz -lpthread -o tiecov_O0
mkdir -p "$OUT"
cp tiebrush tiecov tiebrush_O0 tiecov_O0 hts_tool include/htslib/htsfile "$OUT/"
cp "$REF/tiewrap.py" "$OUT/tiewrap.py"
echo "build_ref: built $(ls "$OUT" | tr '\n' ' ')"
