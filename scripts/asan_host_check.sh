#!/bin/bash
# scripts/asan_host_check.sh -- the HOST code of libqbgpu (table builders, the host restatements behind qbgpu_debug_*_host, the
# tridiagonal / Hermitian eigensolvers, row partitioning, argument checks) under AddressSanitizer + UndefinedBehaviorSanitizer:
# a second build of the library into /tmp (the in-tree libqbgpu.so is not touched), then the CPU tests that call into it.
# No GPU needed.  Last run (end of round 2): 130 tests, no report.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-/tmp/qbgpu_asan}
mkdir -p "$OUT/_build"
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ --expt-relaxed-constexpr -Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined -Xcompiler -fno-omit-frame-pointer -Xcompiler -g"
cd "$ROOT/quantum_basis_b200/csrc"
for f in context spmv sjds sjds_bulk vecops matrix krylov builders peer trlan sectors orbit species dist; do
    ( $NV -c $f.cu -o "$OUT/_build/$f.o" ) &
    if (( $(jobs -r | wc -l) >= 8 )); then wait -n; fi
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined -o "$OUT/libqbgpu.so" "$OUT"/_build/*.o
cat > "$OUT/run.py" <<PY
import sys
sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tests")
import quantum_basis_b200._lib as L
L.LIB_PATH = "$OUT/libqbgpu.so"
import pytest
sys.exit(pytest.main(["-q", "-s", "-p", "no:cacheprovider", "-m", "not gpu", "-k", "not cpp and not sm100a and not every_declared"] + sys.argv[1:]))
PY
cd "$ROOT"
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 \
    python "$OUT/run.py" tests/test_species_cpu.py tests/test_builders_cpu.py tests/test_abi.py tests/test_ckpt_cpu.py > "$OUT/out.txt" 2>&1 || true
tail -2 "$OUT/out.txt"
echo "sanitizer reports: $(grep -c 'AddressSanitizer\|runtime error' "$OUT/out.txt" || true)"
