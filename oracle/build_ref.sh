#!/bin/bash
# Builds the UNMODIFIED reference (fixstars/cuda-bundle-adjustment) for sm_100 into oracle/_ref/
# from the sources where they lie under /root/reference -- nothing is copied into the repo.
# Eigen is not installed in this image: the reference is compiled against the container-only
# stand-in in include/cuba_compat/Eigen (the reference uses Eigen for containers only).
# TEST INFRASTRUCTURE: the result is the GPU parity oracle and bench.py's `--impl reference` arm.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
if [ ! -d "$REF/src" ]; then echo "build_ref: $REF absent - keeping prebuilt files"; exit 0; fi
mkdir -p "$OUT"
build() {  # $1 = output name, $2.. = extra flags
	local out="$1"; shift
	"$NVCC" -O3 -std=c++17 -gencode arch=compute_100,code=sm_100 -lineinfo -w \
		-Xcompiler -fPIC -shared "$@" \
		-I "$REF/include" -I "$REF/src" -I "$HERE/../include/cuba_compat" \
		-o "$OUT/$out" \
		"$REF/src/cuda_bundle_adjustment.cpp" "$REF/src/cuda_linear_solver.cpp" "$REF/src/sparse_block_matrix.cpp" \
		"$REF/src/cuda_block_solver.cu" "$HERE/ref_driver.cpp" \
		-lcusolver -lcusparse
}
if [ ! -f "$OUT/libcuba_ref.so" ] || [ "$HERE/ref_driver.cpp" -nt "$OUT/libcuba_ref.so" ]; then
	build libcuba_ref.so
	echo "built $OUT/libcuba_ref.so"
fi
if [ ! -f "$OUT/libcuba_ref_f32.so" ] || [ "$HERE/ref_driver.cpp" -nt "$OUT/libcuba_ref_f32.so" ]; then
	build libcuba_ref_f32.so -DUSE_FLOAT32
	echo "built $OUT/libcuba_ref_f32.so"
fi
