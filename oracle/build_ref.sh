#!/usr/bin/env bash
# Builds the REFERENCE'S OWN prefilter plugin `rfstudio_render_utils` from its sources where they lie under
# /root/reference (nothing is copied), as a torch extension for sm_100a, into oracle/_ref/ (git-ignored, but it
# travels to the GPU box).  It is used only as a checker: tests/test_prefilter_gpu.py compares both the C oracle
# and the CUDA product path against it.  The reference JIT-builds the same three files with
# torch.utils.cpp_extension.load (rfstudio/graphics/_mesh/_splitsum/_wrap.py:52-72); this is that recipe made
# explicit.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC=/root/reference/rfstudio/graphics/_mesh/_splitsum/c_src
OUT="$HERE/_ref"
[ -d "$SRC" ] || { echo "reference sources not present; keeping any prebuilt $OUT" >&2; exit 0; }
if [ -f "$OUT/rfstudio_render_utils.so" ] && [ -z "$(find "$SRC" -newer "$OUT/rfstudio_render_utils.so" -type f | head -1)" ] && [ -z "${FORCE:-}" ]; then
    echo "up to date: $OUT/rfstudio_render_utils.so"; exit 0
fi
mkdir -p "$OUT/obj"
PY=${PYTHON:-python}
INC=$($PY - <<'PY'
import sysconfig, torch.utils.cpp_extension as e
print(" ".join("-I" + p for p in e.include_paths() + [sysconfig.get_paths()["include"]]))
PY
)
LIBDIR=$($PY -c "import torch, os; print(os.path.join(os.path.dirname(torch.__file__), 'lib'))")
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DNVDR_TORCH -DTORCH_EXTENSION_NAME=rfstudio_render_utils -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1 --expt-relaxed-constexpr -w"
$NVCC $FLAGS $INC -c "$SRC/cubemap.cu" -o "$OUT/obj/cubemap.o"
$NVCC $FLAGS $INC -x cu -c "$SRC/common.cpp" -o "$OUT/obj/common.o"
$NVCC $FLAGS $INC -x cu -c "$SRC/torch_bindings.cpp" -o "$OUT/obj/torch_bindings.o"
$NVCC -shared -o "$OUT/rfstudio_render_utils.so" "$OUT/obj/cubemap.o" "$OUT/obj/common.o" "$OUT/obj/torch_bindings.o" \
    -L"$LIBDIR" -Xlinker -rpath -Xlinker "$LIBDIR" -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python
rm -rf "$OUT/obj"
echo "built $OUT/rfstudio_render_utils.so"
