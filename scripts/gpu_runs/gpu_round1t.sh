#!/bin/bash
mkdir -p gpurun_out
for w in tri21_k10 tri31_k10; do
  timeout -k 5 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"; tail -c 2600 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
