#!/bin/bash
# Device sector assembler: parity tests, then the config-2 / config-5 sized sectors.
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "sector" > gpurun_out/pytest_sectors.log 2>&1; echo "pytest sectors rc=$?"; tail -15 gpurun_out/pytest_sectors.log
for w in heis_chain24_k3 heis_chain28_k1 heis_chain32_k0; do
  timeout -k 5 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"; tail -c 3000 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
