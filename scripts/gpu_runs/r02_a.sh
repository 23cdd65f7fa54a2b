#!/bin/bash
# Round 2, first GPU call: name the failing tests of the formerly xfail-masked module; ncu of the species-order kernels.
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_zz_gpu_species.py -q --runxfail -rA -p no:cacheprovider > gpurun_out/r02_pytest_species.log 2>&1; echo "pytest species rc=$?"
grep -E "^(PASSED|FAILED|ERROR)|passed|failed" gpurun_out/r02_pytest_species.log | tail -45
grep -E "^E  " gpurun_out/r02_pytest_species.log | head -40
for kind in stored matrix_free; do
  timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'spmv_sjds|kron_' -s 2 -c 2 -o gpurun_out/r02_prof_species_$kind \
     python scripts/species_ncu_target.py hubbard4x4 $kind 2 > gpurun_out/r02_ncu_species_$kind.log 2>&1; echo "ncu $kind rc=$?"; tail -3 gpurun_out/r02_ncu_species_$kind.log
done
timeout -k 5 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'spmv_sjds|kron_|perm|native' -c 40 --csv --log-file gpurun_out/r02_launches_species_probe.csv \
   python bench.py --species-probe --workload hubbard4x4 --steps 2 > gpurun_out/r02_species_probe_ncu.log 2>&1; echo "ncu probe rc=$?"
QBGPU_SPECIES_TILE=128 timeout 600 python bench.py --species-probe --workload hubbard4x4 --steps 10 > gpurun_out/r02_species_probe_W128.json 2> gpurun_out/r02_species_probe_W128.err; tail -c 3000 gpurun_out/r02_species_probe_W128.json
ls -la gpurun_out | tail -12
