#!/usr/bin/env bash
# oracle/build_oracle.sh -- builds the two CPU checkers.  TEST INFRASTRUCTURE ONLY.
#
#   oracle/liboracle.so          the C restatement (bm_oracle.c); always built.
#   oracle/_ref/libref_bm.so     the reference's OWN code (kernel1.cl + BoyreMoore.cpp:13-68,
#                                153-190) compiled where it lies under /root/reference through
#                                the macro shim in ref_shim/; built only when /root/reference
#                                exists (i.e. in the authoring container -- the GPU box uses the
#                                prebuilt file, which travels with the snapshot).
#   oracle/_ref/BoyreMoore_ref   the UNMODIFIED BoyreMoore.cpp linked against the stub
#                                ref_shim/CL/cl2.hpp (whole program, reads inputEd.txt /
#                                input1Search.txt / kernel1.cl from its cwd like the original).
# Reference sources are only ever read from /root/reference and piped to the compiler; nothing
# is copied into the repo.  oracle/_ref/ is git-ignored but not gpurun-ignored.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${BMX_REFERENCE_ROOT:-/root/reference}"
CC="${CC:-gcc}"; CXX="${CXX:-g++}"

# No -march=native: the .so files are built here and run on the GPU box's (different) host CPU.
"$CC" -O3 -fPIC -shared -std=c11 -D_GNU_SOURCE -Wall -Wextra -o "$here/liboracle.so" "$here/bm_oracle.c" -lpthread
echo "built $here/liboracle.so"

cpp_src="$ref/BoyreMoore/BoyreMoore/BoyreMoore.cpp"
if [[ -f "$cpp_src" && -f "$ref/BoyreMoore/x64/Debug/kernel1.cl" ]]; then
    mkdir -p "$here/_ref"
    {
        cat "$here/ref_shim/head.inc"
        sed -n '13,68p' "$cpp_src"
        cat "$here/ref_shim/mid.inc"
        sed -n '153,190p' "$cpp_src"
        cat "$here/ref_shim/tail.inc"
    } | "$CXX" -O3 -fPIC -shared -std=c++17 -w -DREF_KERNEL_PATH="\"$ref/BoyreMoore/x64/Debug/kernel1.cl\"" -x c++ - -o "$here/_ref/libref_bm.so" -lpthread
    echo "built $here/_ref/libref_bm.so (reference code compiled from $ref)"
    "$CXX" -O3 -w -std=c++17 -I"$here/ref_shim" "$cpp_src" -o "$here/_ref/BoyreMoore_ref"
    echo "built $here/_ref/BoyreMoore_ref (unmodified BoyreMoore.cpp + stub CL/cl2.hpp)"
else
    echo "reference tree not found at $ref: keeping any prebuilt oracle/_ref/ as is"
fi
