#!/bin/bash
# Third GPU pass: kernel-variant sweep (sliced-jagged kernel; launch bounds / unroll / cache policies), ncu DRAM bytes.
mkdir -p gpurun_out
timeout -k 5 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for t in tests/test_gpu_parity.py::test_double_precision_real_matrix_path tests/test_gpu_parity.py::test_large_generated_matrix_properties tests/test_gpu_parity.py::test_vec_randomize_is_the_reference_sequence tests/test_gpu_parity.py::test_sliced_jagged_layout_gives_the_same_product tests/test_gpu_parity.py::test_multmv_matches_reference_product tests/test_gpu_parity.py::test_lanczos_E0_matches_reference_and_published_golden; do
  timeout -k 5 150 python -m pytest $t -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pt.log 2>&1; echo "rc=$? $t"; grep -E "^E  |passed|failed" gpurun_out/pt.log | head -5
done
timeout -k 5 300 python scripts/kbench.py heis_chain28 --csr > gpurun_out/kbench_heis28.txt 2>&1; echo "kbench heis28 rc=$?"; cat gpurun_out/kbench_heis28.txt
timeout -k 5 300 python scripts/kbench.py heis_chain28 --real > gpurun_out/kbench_heis28_real.txt 2>&1; echo "kbench heis28 real rc=$?"; cat gpurun_out/kbench_heis28_real.txt
timeout -k 5 420 python scripts/kbench.py hubbard4x4 --csr > gpurun_out/kbench_hubbard4x4.txt 2>&1; echo "kbench hubbard rc=$?"; cat gpurun_out/kbench_hubbard4x4.txt
timeout -k 5 420 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_evict_first_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_miss.sum --clock-control none -k regex:spmv_sjds -c 12 --csv --log-file gpurun_out/ncu_hubbard_dram.csv python scripts/kbench.py hubbard4x4 --ids 0,4,16 --reps 1 > gpurun_out/ncu_hubbard.log 2>&1; echo "ncu hubbard rc=$?"
ls -la gpurun_out | head -40
