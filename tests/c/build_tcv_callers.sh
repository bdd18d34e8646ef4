#!/bin/sh
# Builds tests/c/tcv_caller.c four ways (the variants that need the reference's header are built only where
# /root/reference exists; the binaries are git-ignored and travel to the GPU box):
#   tcv_caller           include/tcvideo.h + libtcvgpu + libacgpu            -- this repo's tcv_* interface
#   tcv_caller_ref       reference header + reference libtcvideo over the plain-C aclib
#   tcv_caller_ref_sse2  reference header + reference libtcvideo over the SSE2 aclib (stock transcode)
#   tcv_caller_legacy    reference header + reference libtcvideo over libacgpu  -- the unmodified per-row ac_* calls
# `tests/c/<variant> time N` then times N frames of 1080p linear blend through each.
set -e
cd "$(dirname "$0")/../.."
ROOT=$PWD
PKG=$ROOT/transcode-tcforge_b200
REF=${REF:-/root/reference}
gcc -std=gnu99 -O1 -Wall -Werror -Iinclude -o tests/c/tcv_caller tests/c/tcv_caller.c -L"$PKG" -ltcvgpu -lacgpu -Wl,-rpath,"$PKG"
if [ -d "$REF/libtcvideo" ]; then
    for v in ref:libtcv_ref.so ref_sse2:libtcv_ref_sse2.so legacy:libtcv_over_acgpu.so; do
        name=${v%%:*}; lib=${v##*:}
        [ -f "oracle/_ref/$lib" ] || continue
        gcc -std=gnu99 -O1 -Wall -DHAVE_CONFIG_H -Ioracle/refcfg -I"$REF/libtcvideo" -I"$REF" -o "tests/c/tcv_caller_$name" tests/c/tcv_caller.c \
            -Loracle/_ref -l:"$lib" -Wl,-rpath,"$ROOT/oracle/_ref" -Wl,-rpath,"$PKG" -L"$PKG" $( [ "$name" = legacy ] && echo -lacgpu )
    done
fi
ls tests/c/
