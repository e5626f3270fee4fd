#!/bin/bash
# ORACLE build recipe for oracle/_ref/ — the reference's own FSR.cl (OpenCL C) compiled for the CPU.
#
#   usage: build_ref.sh [/root/reference]        (silently does nothing when the reference tree is absent:
#                                                  the GPU box only uses the prebuilt oracle/_ref/*.so)
#
# The kernel source is read WHERE IT LIES (LiveVisionKit/Functions/OpenCL/Sources/FSR.cl) and piped into g++; no copy
# of it is written anywhere.  Two in-flight, purely syntactic adaptations make OpenCL C acceptable to a C++ compiler:
#   1. line 1 of the file, `R"(`, is dropped (the file is a C++ raw string literal body for Kernels.hpp; its closing
#      `)"` and the `)" R"(` splices sit behind `//` and stay comments);
#   2. OpenCL vector literals `(float4)(a, b, c, d)` become constructor calls `float4(a, b, c, d)` (in C++ the former is
#      a cast of a comma expression).  One regex, applied to type names only.
# Everything else (vector types, as_/convert_/vload/vstore, work-item ids, built-ins) is supplied by opencl_c_shim.hpp.
# The text is compiled twice into one library: namespace fsr_bgr (no flags) and fsr_yuv (-D YUV_INPUT), Image.cpp:37-38.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../_ref"
CL="$REF/LiveVisionKit/Functions/OpenCL/Sources/FSR.cl"
[ -f "$CL" ] || { echo "build_ref.sh: $CL not found - keeping any prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$OUT"

kernel_text() { tail -n +2 "$CL" | sed -E 's/\((float|int|uint|uchar)(2|3|4|8|16)?\)\(/\1\2(/g'; }
unit() {
    echo '#include "opencl_c_shim.hpp"'
    echo 'namespace fsr_bgr {'; kernel_text; echo '}'
    echo '#define YUV_INPUT'
    echo 'namespace fsr_yuv {'; kernel_text; echo '}'
    echo '#undef YUV_INPUT'
    echo '#line 1 "fsr_cl_driver.cpp"'
    cat "$HERE/fsr_cl_driver.cpp"
}
FMA_FLAG=""
grep -q -m1 ' fma ' /proc/cpuinfo && FMA_FLAG="-mfma"
COMMON="-x c++ -std=gnu++17 -O2 -fno-fast-math -fPIC -shared -pthread -I$HERE -Wno-unused-function"
newer() { [ -f "$1" ] && [ "$1" -nt "$CL" ] && [ "$1" -nt "$HERE/opencl_c_shim.hpp" ] && [ "$1" -nt "$HERE/fsr_cl_driver.cpp" ] && [ "$1" -nt "$HERE/build_ref.sh" ]; }
newer "$OUT/libfsrcl_ref_strict.so"   || unit | g++ $COMMON -ffp-contract=off  $FMA_FLAG -o "$OUT/libfsrcl_ref_strict.so" -
if [ -n "$FMA_FLAG" ]; then
newer "$OUT/libfsrcl_ref_contract.so" || unit | g++ $COMMON -ffp-contract=fast $FMA_FLAG -o "$OUT/libfsrcl_ref_contract.so" -
fi
