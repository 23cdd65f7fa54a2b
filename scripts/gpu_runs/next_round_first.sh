#!/bin/bash
# First GPU call of the next round: the species-order handles (written after round 1's GPU budget was spent).
#   gpurun --timeout 900 -- 'bash scripts/gpu_runs/next_round_first.sh'
set -x
mkdir -p gpurun_out
# 1. correctness: the xfail-marked module, run strictly (a failure here is a failure)
python -m pytest tests/test_zz_gpu_species.py -q -x --runxfail -p no:cacheprovider > gpurun_out/pytest_species.log 2>&1; echo "pytest species rc=$?" >> gpurun_out/pytest_species.log
tail -5 gpurun_out/pytest_species.log
# 1b. if anything above failed: memcheck on the smallest case pinpoints an out-of-bounds access
if ! grep -q "rc=0" gpurun_out/pytest_species.log; then
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_zz_gpu_species.py -q -x --runxfail -p no:cacheprovider \
    -k "product_in_the_reference_order and hub2x2" > gpurun_out/memcheck_species.log 2>&1; tail -30 gpurun_out/memcheck_species.log
fi
# 2. timings on BASELINE config 3, tile sweep (the probe itself sweeps the three pass-1 variants of the matrix-free product)
for W in 64 128 256 512; do
  QBGPU_SPECIES_TILE=$W timeout 600 python bench.py --species-probe --workload hubbard4x4 --steps 10 > gpurun_out/species_probe_W$W.json 2> gpurun_out/species_probe_W$W.err
  tail -c 1500 gpurun_out/species_probe_W$W.json
done
# 3. the term-coded replay kernel v3 left over from round 1
QBGPU_TERMS_KERNEL=3 timeout 300 python scripts/terms_check.py > gpurun_out/terms_check_v3.txt 2>&1; tail -3 gpurun_out/terms_check_v3.txt
# 4. the C++ examples through the adaptor (reference asserts: chain L = 16 all momenta; Hubbard 4x2 E0, all three handle kinds)
g++ -std=c++17 -O2 -I include examples/square_fermi_hubbard.cc -L quantum_basis_b200 -lqbgpu -Wl,-rpath,$PWD/quantum_basis_b200 -o /tmp/square_fermi_hubbard \
  && /tmp/square_fermi_hubbard 4 2 4 4 > gpurun_out/example_hubbard.txt 2>&1; echo "example rc=$?" >> gpurun_out/example_hubbard.txt; cat gpurun_out/example_hubbard.txt
# 5. the full bench line on the species-order handles (headline contract: reference-order vectors, e2e, Lanczos, clocks)
for lay in species species-matfree; do
  timeout 900 python bench.py --layout $lay --no-species --no-cpu > gpurun_out/bench_layout_$lay.json 2> gpurun_out/bench_layout_$lay.err; tail -c 1200 gpurun_out/bench_layout_$lay.json
done
