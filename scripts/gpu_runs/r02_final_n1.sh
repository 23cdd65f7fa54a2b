#!/bin/bash
# Round 2 record pass (1 GPU): smoke, all GPU tests, the bench line, the reference arm, ncu launch list of the bench command,
# ncu --set full of one reference-order MultMv (way in, pass 1, pass 2, way out).
mkdir -p gpurun_out
timeout -k 5 200 python __graft_entry__.py smoke > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02f_smoke.log
timeout -k 5 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rs > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02f_pytest_gpu.log
grep -E "^E  |^FAILED" gpurun_out/r02f_pytest_gpu.log | head -20
timeout -k 5 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02f_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_n1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','roofline','e2e','parity_sampled','lanczos','real_vectors','cpu_baseline','gpu_launches','clocks'):
    print(k, json.dumps(d.get(k))[:600])
so=d.get('species_order',{})
for k in ('stored','matrix_free'):
    print(k, {kk:(round(v['ms_per_product'],3) if isinstance(v,dict) and 'ms_per_product' in v else v) for kk,v in so.get(k,{}).items() if kk!='pass1_variants_complex_ms'})
print('ordinary_ms', so.get('ordinary_ms'))
PY
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench_hubbard4x4.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-lanczos --no-species > gpurun_out/r02f_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'native|spmv_sjds|sjds_block' -s 4 -c 4 -o gpurun_out/r02f_prof_mv_reference_order python scripts/mv_ncu_target.py hubbard4x4 3 > gpurun_out/r02f_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout -k 5 600 python bench.py --workload hubbard4x3 --steps 20 --warmup 5 --no-species > gpurun_out/r02f_bench_hubbard4x3.json 2> /dev/null; timeout -k 5 300 python bench.py --impl reference --workload hubbard4x3 --steps 20 --warmup 5 > gpurun_out/r02f_bench_reference_hubbard4x3.json 2>/dev/null; tail -c 400 gpurun_out/r02f_bench_reference_hubbard4x3.json
for w in heis_chain32_k0 tri31_k10 heis_chain28_k1 heis_chain20; do timeout -k 5 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/r02f_bench_$w.json 2> /dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02f_bench_$w.json').read().strip().splitlines()[-1]); print('$w', round(d['ms_per_step'],4), round(d['roofline']['frac'],3), d.get('lanczos',{}).get('E0'), d.get('lanczos',{}).get('iters_per_s'))
PY
done
