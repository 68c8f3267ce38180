#!/usr/bin/env bash
# Test infrastructure only: builds the REFERENCE's own CUDA kernels (unmodified sources, read where
# they lie under /root/reference) for sm_100a into oracle/_ref/libpdr_ref_cuda.so, so that the GPU
# parity tests can compare our kernels bit-for-bit with the reference's on the B200 box.
# Nothing is copied into the repo; outputs go to oracle/_ref/ only (git-ignored, travels via gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${PDR_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
mkdir -p "$OUT/obj"
if [ ! -d "$REF" ]; then echo "reference tree $REF not present; keeping prebuilt $OUT" ; exit 0; fi
PY="${PYTHON:-python}"
TORCH_DIR="$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))')"
PY_INC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
EXT="$REF/pointnet2_ops_lib/pointnet2_ops/_ext-src"
EMD="$REF/PytorchEMD/cuda"
CH3="$REF/pointnet2/models/pvd/metrics/ChamferDistancePytorch/chamfer3D"
INC=(-I"$EXT/include" -I"$HERE/compat" -I"$TORCH_DIR/include" -I"$TORCH_DIR/include/torch/csrc/api/include" -I"$PY_INC")
NVFLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Xcompiler -fPIC -DTORCH_EXTENSION_NAME=pdr_ref -D_GLIBCXX_USE_CXX11_ABI=1 -w)
build() { # src obj
  if [ ! -f "$2" ] || [ "$1" -nt "$2" ]; then echo "nvcc $1"; nvcc -c "${NVFLAGS[@]}" "${INC[@]}" "$1" -o "$2"; fi
}
pids=()
build "$EXT/src/sampling_gpu.cu"     "$OUT/obj/sampling_gpu.o" & pids+=($!)
build "$EXT/src/ball_query_gpu.cu"   "$OUT/obj/ball_query_gpu.o" & pids+=($!)
build "$EXT/src/group_points_gpu.cu" "$OUT/obj/group_points_gpu.o" & pids+=($!)
build "$EXT/src/interpolate_gpu.cu"  "$OUT/obj/interpolate_gpu.o" & pids+=($!)
build "$EMD/emd_kernel.cu"           "$OUT/obj/emd_kernel.o" & pids+=($!)
build "$CH3/chamfer3D.cu"            "$OUT/obj/chamfer3D.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
if [ ! -f "$OUT/obj/ref_shim.o" ] || [ "$HERE/ref_shim.cpp" -nt "$OUT/obj/ref_shim.o" ]; then
  echo "g++ ref_shim.cpp"
  g++ -c -O2 -std=c++17 -fPIC -w -D_GLIBCXX_USE_CXX11_ABI=1 "${INC[@]}" -I/usr/local/cuda/include "$HERE/ref_shim.cpp" -o "$OUT/obj/ref_shim.o"
fi
g++ -shared -o "$OUT/libpdr_ref_cuda.so" "$OUT"/obj/*.o \
  -L"$TORCH_DIR/lib" -Wl,-rpath,"$TORCH_DIR/lib" -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch \
  -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -lcudart
echo "built $OUT/libpdr_ref_cuda.so"
# The reference's own Python for the literal drop-in test on the GPU box (tests/test_dropin_gpu.py imports
# pointnet2/completion_eval.py ITSELF and runs it on libpdr_b200.so).  Staged UNMODIFIED into oracle/_ref/pyref/ only:
# git-ignored test infrastructure that travels with the snapshot exactly like the .so above, never part of the repo.
PYREF="$OUT/pyref"
rm -rf "$PYREF"; mkdir -p "$PYREF/pointnet2" "$PYREF/pointnet2_ops_lib"
( cd "$REF/pointnet2" && find . -maxdepth 3 -name "*.py" -not -path "./models/pvd/*" -not -path "./exp_configs/*" -print0 | tar --null -T - -cf - ) | tar -xf - -C "$PYREF/pointnet2"
( cd "$REF/pointnet2_ops_lib" && find pointnet2_ops -maxdepth 1 -name "*.py" -print0 | tar --null -T - -cf - ) | tar -xf - -C "$PYREF/pointnet2_ops_lib"
echo "staged $(find "$PYREF" -name '*.py' | wc -l) reference .py files into $PYREF"
