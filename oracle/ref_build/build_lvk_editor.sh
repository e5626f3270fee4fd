#!/bin/bash
# Drop-in proof, build recipe for oracle/_ref/lvk-editor — the reference's own VideoEditor CLI (Modules/VideoEditor:
# Application.cpp, VideoProcessor.cpp, VideoIOConfiguration.cpp, ConsoleLogger.cpp and the header-only Option/Filter
# parsers) compiled UNCHANGED, where the sources lie, against this repo's boundary instead of the LiveVisionKit library:
#   * `#include <LiveVisionKit.hpp>` resolves to a generated redirect: livevisionkit_b200/compat/lvk/lvk.hpp (built with
#     the reference's `struct VideoFrame : cv::UMat`) plus the reference's own Logger / CSVLogger (support classes that
#     are not on the accelerated path: taken in place from the reference tree, as a maintainer would keep them);
#   * OpenCV is the mock under tests/cpp/mock_opencv (raw-clip VideoCapture / VideoWriter, no-op highgui);
#   * it links liblvkb200.so (rpath relative to the binary, so the prebuilt file runs on the GPU box).
#
#   usage: build_lvk_editor.sh [/root/reference]     (does nothing when the reference tree is absent)
#
# Nothing of the reference is copied into the repo: the only output is the binary under oracle/_ref/ (git-ignored).
# Test infrastructure: tests/test_compat_cpu.py builds and runs its manual; tests/test_compat_gpu.py runs a clip through it.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="${LVK_EDITOR_OUT:-$ROOT/oracle/_ref}"
ED="$REF/Modules/VideoEditor"
[ -f "$ED/Application.cpp" ] || { echo "build_lvk_editor.sh: $ED not found - keeping any prebuilt oracle/_ref/lvk-editor" >&2; exit 0; }
[ -f "$ROOT/livevisionkit_b200/liblvkb200.so" ] || { echo "build_lvk_editor.sh: build liblvkb200.so first" >&2; exit 1; }
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cat > "$TMP/LiveVisionKit.hpp" <<HDR
#pragma once
#define LVK_COMPAT_USE_OPENCV
#include <opencv2/opencv.hpp>
#include <opencv2/core/ocl.hpp>
#include "$ROOT/livevisionkit_b200/compat/lvk/lvk.hpp"
#include "$REF/LiveVisionKit/Logging/Logger.hpp"
#include "$REF/LiveVisionKit/Logging/CSVLogger.hpp"
HDR
# Logger.tpp asks for "Directives.hpp" (the assertion macros): lvk.hpp carries them
printf '#pragma once\n#include "%s/livevisionkit_b200/compat/lvk/lvk.hpp"\n' "$ROOT" > "$TMP/Directives.hpp"
FLAGS="-std=c++20 -O1 -Werror -Wno-unused-result -I$TMP -I$ROOT/tests/cpp/mock_opencv -I$ED"
OBJS=""
for src in "$ED/Application.cpp" "$ED/VideoProcessor.cpp" "$ED/VideoIOConfiguration.cpp" "$ED/ConsoleLogger.cpp" \
           "$REF/LiveVisionKit/Logging/CSVLogger.cpp"; do
    obj="$TMP/$(basename "$src" .cpp).o"
    g++ $FLAGS -c "$src" -o "$obj"
    OBJS="$OBJS $obj"
done
g++ $OBJS -o "$OUT/lvk-editor" -L"$ROOT/livevisionkit_b200" -l:liblvkb200.so '-Wl,-rpath,$ORIGIN/../../livevisionkit_b200'
echo "built $OUT/lvk-editor"
