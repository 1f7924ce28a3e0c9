#!/usr/bin/env bash
# shim/build_shim.sh -- builds the REFERENCE ENCODER WITH THE B200 SHIM: a scratch copy of /root/reference/Source (deleted afterwards) + shim/b200.patch
# (two files of Source/Lib/Codec, nothing under Source/API) + shim/EbB200Shim.{c,h}, compiled with the same per-directory flags as oracle/build_ref.sh
# (AVX2 tree as-is) into
#   oracle/_ref/libsvtref_b200.so      the patched library
#   oracle/_ref/SvtHevcEncApp_b200     the reference's own CLI linked against it
# The shim dlopen()s svt-hevc_b200/libsvthevc_b200.so at EbInitEncoder when SVT_HEVC_B200=1; with the variable unset the binary runs the reference's own
# code path.  tests/test_gpu_shim.py encodes the same clip both ways and compares the bitstream and reconstruction byte for byte.
set -euo pipefail
REF=${SVT_REFERENCE_ROOT:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/../oracle/_ref
JOBS=${JOBS:-$(nproc)}
if [ ! -d "$REF/Source/Lib/Codec" ]; then
  echo "build_shim.sh: $REF not present; keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
STAMP="$OUT/.stamp_shim"
if [ -f "$STAMP" ] && [ -f "$OUT/libsvtref_b200.so" ] && [ -f "$OUT/SvtHevcEncApp_b200" ] \
   && [ -z "$(find "$HERE" "$HERE/../include" -maxdepth 1 -type f -newer "$STAMP")" ]; then
  echo "build_shim.sh: up to date"; exit 0
fi
SCR=$(mktemp -d)
cp -r "$REF/Source" "$SCR/Source"
chmod -R u+w "$SCR/Source"
# the public API must stay untouched: remember its checksum, verify after patching
API_BEFORE=$(cd "$SCR/Source/API" && cat * | sha256sum)
patch -s -p1 -d "$SCR" < "$HERE/b200.patch"
cp "$HERE/EbB200Shim.c" "$HERE/EbB200Shim.h" "$SCR/Source/Lib/Codec/"
API_AFTER=$(cd "$SCR/Source/API" && cat * | sha256sum)
[ "$API_BEFORE" = "$API_AFTER" ] || { echo "build_shim.sh: Source/API changed" >&2; exit 1; }
SRC=$SCR/Source
OBJ=$(mktemp -d)
sed -e 's/@SVT_HEVC_VERSION_MAJOR@/1/' -e 's/@SVT_HEVC_VERSION_MINOR@/5/' -e 's/@SVT_HEVC_VERSION_PATCHLEVEL@/1/' "$SRC/API/EbApiVersion.h.in" > "$OBJ/EbApiVersion.h"
INC="-I$OBJ -I$SRC/API -I$SRC/Lib/Codec -I$SRC/Lib/C_DEFAULT -I$SRC/Lib/ASM_SSE2 -I$SRC/Lib/ASM_SSSE3 -I$SRC/Lib/ASM_SSE4_1 -I$SRC/Lib/ASM_AVX2 -I$HERE/../include"
BASE="-O2 -std=gnu99 -fPIC -w -D_GNU_SOURCE $INC"
: > "$OBJ/cmds"
for d in Codec C_DEFAULT; do
  for f in "$SRC"/Lib/$d/*.c; do echo "gcc $BASE -c $f -o $OBJ/${d}_$(basename "${f%.c}").o" >> "$OBJ/cmds"; done
done
for d in ASM_SSE2 ASM_SSSE3 ASM_SSE4_1; do
  for f in "$SRC"/Lib/$d/*.c; do echo "gcc $BASE -msse4.1 -c $f -o $OBJ/${d}_$(basename "${f%.c}").o" >> "$OBJ/cmds"; done
done
for f in "$SRC"/Lib/ASM_AVX2/*.c; do echo "gcc $BASE -mavx2 -c $f -o $OBJ/ASM_AVX2_$(basename "${f%.c}").o" >> "$OBJ/cmds"; done
echo "gcc $BASE -c $HERE/../oracle/yasm_shim.c -o $OBJ/yasm_shim.o" >> "$OBJ/cmds"
xargs -P "$JOBS" -I{} sh -c '{}' < "$OBJ/cmds"
gcc -shared -o "$OUT/libsvtref_b200.so" "$OBJ"/*.o -lpthread -lm -ldl
gcc -O2 -std=gnu99 -w -D_GNU_SOURCE -I"$OBJ" -I"$SRC/API" "$SRC"/App/*.c -o "$OUT/SvtHevcEncApp_b200" -L"$OUT" -lsvtref_b200 -Wl,-rpath,'$ORIGIN' -lpthread -lm -ldl
rm -rf "$OBJ" "$SCR"
touch "$STAMP"
ls -la "$OUT/libsvtref_b200.so" "$OUT/SvtHevcEncApp_b200"
