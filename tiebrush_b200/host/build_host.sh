#!/usr/bin/env bash
# Builds the host side of the drop-in: tiebrush_gpu / tiecov_gpu = the reference's own command lines with the hot loop
# replaced by libtiebrush_b200.so (see tiebrush_gpu_main.cpp / tiecov_gpu_main.cpp). The reference sources are compiled
# FROM WHERE THEY LIE under $REF (nothing is copied into the repo); htslib is built on a scratch copy exactly as
# oracle/build_ref.sh does. Outputs: tiebrush_b200/host/_build/{tiebrush_gpu,tiecov_gpu} (git-ignored; they travel to
# the GPU box with the snapshot, where /root/reference does not exist and this script is a no-op).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/_build"
if [ ! -d "$REF/src" ]; then
  echo "build_host: $REF not present (GPU box?) - using prebuilt $OUT if any" >&2
  exit 0
fi
if [ -x "$OUT/tiebrush_gpu" ] && [ -x "$OUT/tiecov_gpu" ] && [ "${FORCE:-0}" != 1 ] \
   && [ "$OUT/tiebrush_gpu" -nt "$HERE/tiebrush_gpu_main.cpp" ] && [ "$OUT/tiecov_gpu" -nt "$HERE/tiecov_gpu_main.cpp" ] \
   && [ "$OUT/tiebrush_gpu" -nt "$ROOT/include/tiebrush_b200.h" ]; then
  exit 0
fi
W=$(mktemp -d /tmp/hostbuild.XXXXXX)
trap 'rm -rf "$W"' EXIT
mkdir -p "$OUT"
if [ ! -f "$OUT/libhts.a" ]; then
  cp -r "$REF/include/htslib" "$W/htslib"
  chmod -R u+w "$W/htslib"
  ( cd "$W/htslib"
    printf '#ifndef _XOPEN_SOURCE\n#define _XOPEN_SOURCE 600\n#endif\n#define HAVE_DRAND48 1\n' > config.h
    echo '#define HTSCODECS_VERSION_TEXT "1.5.0"' > htscodecs/htscodecs/version.h
    make -j"$(nproc)" lib-static NONCONFIGURE_OBJS= >/dev/null 2>&1 )
  cp "$W/htslib/libhts.a" "$OUT/libhts.a"
fi
mkdir -p "$W/stub/libBigWig"
cat > "$W/stub/libBigWig/bigWig.h" <<'EOS'
#pragma once
#include <stdint.h>
#include <stddef.h>
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
SRC="$REF/src/GSam.cpp $REF/src/tmerge.cpp $REF/include/gclib/GStr.cpp $REF/include/gclib/GArgs.cpp $REF/include/gclib/GBase.cpp"
CXXF="-std=c++11 -fpermissive -w -DNOCURL=1 -O2 -I$REF -I$REF/src -I$REF/include -I$REF/include/htslib -I$ROOT/include"
LINK="$OUT/libhts.a -L$ROOT/tiebrush_b200 -ltiebrush_b200 -Wl,-rpath,\$ORIGIN/../.. -lz -lpthread"
g++ $CXXF "$HERE/tiebrush_gpu_main.cpp" $SRC $LINK -o "$OUT/tiebrush_gpu"
g++ $CXXF -I"$W/stub" "$HERE/tiecov_gpu_main.cpp" $SRC $LINK -o "$OUT/tiecov_gpu"
echo "build_host: built $(ls "$OUT" | tr '\n' ' ')"
