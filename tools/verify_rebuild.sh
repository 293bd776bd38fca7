#!/bin/sh
# Rebuilds the library from scratch in a scratch directory and compares the SASS of every kernel with the library in the
# tree (no GPU needed).  The anonymous-namespace hash nvcc derives from the source path is the only expected difference.
#   tools/verify_rebuild.sh [scratch dir]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=${1:-/tmp/fanlin_rebuild}
rm -rf "$T" && mkdir -p "$T/fanlin-rs_b200" "$T/include"
cp -r "$ROOT/fanlin-rs_b200/csrc" "$T/fanlin-rs_b200/" && rm -rf "$T"/fanlin-rs_b200/csrc/build*
cp "$ROOT"/include/*.h "$T/include/"
make -s -j8 -C "$T/fanlin-rs_b200/csrc" OUT="$T/libfanlin_device.so" > "$T/make.log" 2>&1
norm() { cuobjdump -sass "$1" | grep -v '^Fatbin\|^=====\|//##' | sed 's/_GLOBAL__N__[0-9a-f]*_/_GLOBAL__N__X_/'; }
norm "$T/libfanlin_device.so" > "$T/rebuilt.sass"
norm "$ROOT/fanlin-rs_b200/libfanlin_device.so" > "$T/tree.sass"
if cmp -s "$T/rebuilt.sass" "$T/tree.sass"; then
    echo "SASS identical: $(grep -c 'Function :' "$T/tree.sass") kernels, $(wc -l < "$T/tree.sass") lines"
else
    echo "SASS differs:"; diff "$T/rebuilt.sass" "$T/tree.sass" | head -20; exit 1
fi
