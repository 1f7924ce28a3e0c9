#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the UNMODIFIED reference sources where they lie under /root/reference into
# oracle/_ref/ (git-ignored, travels to the GPU box with gpurun):
#   libsvtref_avx2.so    tree as-is (NON_AVX512_SUPPORT defined, EbDefinitions.h:18)
#   libsvtref_avx512.so  scratch copy (in $TMPDIR, deleted afterwards) with that one line
#                        commented so the reference's AVX-512 kernels exist/are dispatched
#   SvtHevcEncApp_avx512 the reference CLI linked against the avx512 objects (whole-encoder
#                        fps, reported for context by bench.py --impl reference)
# No version script is applied, so every leaf kernel (SadLoopKernel, Transform32x32Estimate,
# Luma4SampleEdgeDLFCore, ...) is callable through dlsym/ctypes.
# Per-directory ISA flags follow the reference's own CMake files
# (ASM_SSE2/CMakeLists.txt:15-16, ASM_SSSE3/CMakeLists.txt:18-19, ASM_SSE4_1, ASM_AVX2/CMakeLists.txt:19).
# The reference's own build system is NOT run (it needs yasm, which this image lacks).
set -euo pipefail
REF=${SVT_REFERENCE_ROOT:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
JOBS=${JOBS:-$(nproc)}
if [ ! -d "$REF/Source/Lib/Codec" ]; then
  echo "build_ref.sh: $REF not present; keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
STAMP="$OUT/.stamp"
if [ -f "$STAMP" ] && [ -f "$OUT/libsvtref_avx2.so" ] && [ -f "$OUT/libsvtref_avx512.so" ] \
   && [ -f "$OUT/SvtHevcEncApp_avx512" ] && [ "$HERE/build_ref.sh" -ot "$STAMP" ] && [ "$HERE/yasm_shim.c" -ot "$STAMP" ] \
   && [ -z "$(find "$HERE" -maxdepth 1 -name 'ref*_harness.c' -newer "$STAMP")" ]; then
  echo "build_ref.sh: up to date"; exit 0
fi

build_variant() {  # $1 = variant name, $2 = source root
  local V=$1 SRC=$2 OBJ
  OBJ=$(mktemp -d)
  # generated API version header (CMakeLists.txt:51); v1.5.1
  sed -e 's/@SVT_HEVC_VERSION_MAJOR@/1/' -e 's/@SVT_HEVC_VERSION_MINOR@/5/' -e 's/@SVT_HEVC_VERSION_PATCHLEVEL@/1/' \
      "$SRC/API/EbApiVersion.h.in" > "$OBJ/EbApiVersion.h"
  local INC="-I$OBJ -I$SRC/API -I$SRC/Lib/Codec -I$SRC/Lib/C_DEFAULT -I$SRC/Lib/ASM_SSE2 -I$SRC/Lib/ASM_SSSE3 -I$SRC/Lib/ASM_SSE4_1 -I$SRC/Lib/ASM_AVX2"
  local BASE="-O2 -std=gnu99 -fPIC -w -D_GNU_SOURCE $INC"
  : > "$OBJ/cmds"
  for d in Codec C_DEFAULT; do
    for f in "$SRC"/Lib/$d/*.c; do echo "gcc $BASE -c $f -o $OBJ/${d}_$(basename "${f%.c}").o" >> "$OBJ/cmds"; done
  done
  for d in ASM_SSE2 ASM_SSSE3 ASM_SSE4_1; do
    for f in "$SRC"/Lib/$d/*.c; do echo "gcc $BASE -msse4.1 -c $f -o $OBJ/${d}_$(basename "${f%.c}").o" >> "$OBJ/cmds"; done
  done
  for f in "$SRC"/Lib/ASM_AVX2/*.c; do echo "gcc $BASE -mavx2 -c $f -o $OBJ/ASM_AVX2_$(basename "${f%.c}").o" >> "$OBJ/cmds"; done
  echo "gcc $BASE -c $HERE/yasm_shim.c -o $OBJ/yasm_shim.o" >> "$OBJ/cmds"
  # harnesses: our own drivers around unmodified reference functions (include the reference headers in place)
  for f in "$HERE"/ref_*_harness.c; do
    [ -f "$f" ] && echo "gcc $BASE -I$HERE/../include -c $f -o $OBJ/$(basename "${f%.c}").o" >> "$OBJ/cmds"
  done
  xargs -P "$JOBS" -I{} sh -c '{}' < "$OBJ/cmds"
  gcc -shared -o "$OUT/libsvtref_$V.so" "$OBJ"/*.o -lpthread -lm
  # harnesses that #include a reference .c file (to reach its static functions) live in their own shared object so that the
  # duplicate external symbols of that file never meet the originals at link time
  for f in "$HERE"/refinc_*_harness.c; do
    [ -f "$f" ] || continue
    n=$(basename "${f%_harness.c}"); n=${n#refinc_}
    gcc $BASE -I"$HERE/../include" -shared "$f" -o "$OUT/libsvtref_${n}_$V.so" -L"$OUT" -lsvtref_$V -Wl,-rpath,'$ORIGIN' -lpthread -lm
  done
  if [ "$V" = avx512 ]; then
    gcc -O2 -std=gnu99 -w -D_GNU_SOURCE -I"$OBJ" -I"$SRC/API" "$SRC"/App/*.c -o "$OUT/SvtHevcEncApp_avx512" \
        -L"$OUT" -lsvtref_avx512 -Wl,-rpath,'$ORIGIN' -lpthread -lm
  fi
  rm -rf "$OBJ"
}

build_variant avx2 "$REF/Source"

SCR=$(mktemp -d)
cp -r "$REF/Source" "$SCR/Source"
chmod -R u+w "$SCR/Source"
sed -i 's|^#define NON_AVX512_SUPPORT|//#define NON_AVX512_SUPPORT|' "$SCR/Source/Lib/Codec/EbDefinitions.h"
build_variant avx512 "$SCR/Source"
rm -rf "$SCR"
touch "$STAMP"
ls -la "$OUT"
