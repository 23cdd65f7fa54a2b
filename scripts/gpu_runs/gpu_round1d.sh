#!/bin/bash
# Fourth GPU pass: second kernel sweep (per-entry gather policy, launch-bound points).
mkdir -p gpurun_out
IDS=0,20,31,33,34,35,36,37,38,39,40,41,42,43,44
timeout -k 5 420 python scripts/kbench.py hubbard4x4 --ids $IDS --far 17,19,21,23 > gpurun_out/kbench2_hubbard4x4.txt 2>&1; echo "kbench hubbard rc=$?"; cat gpurun_out/kbench2_hubbard4x4.txt
timeout -k 5 300 python scripts/kbench.py heis_chain28 --ids $IDS --far 17,19,21,23 > gpurun_out/kbench2_heis28.txt 2>&1; echo "kbench heis28 rc=$?"; cat gpurun_out/kbench2_heis28.txt
timeout -k 5 300 python scripts/kbench.py heis_chain28 --real --ids $IDS --far 19,21,23 > gpurun_out/kbench2_heis28_real.txt 2>&1; echo "kbench heis28 real rc=$?"; cat gpurun_out/kbench2_heis28_real.txt
timeout -k 5 420 python scripts/kbench.py hubbard4x4 --real --ids $IDS --far 19,21,23 > gpurun_out/kbench2_hubbard4x4_real.txt 2>&1; echo "kbench hubbard real rc=$?"; cat gpurun_out/kbench2_hubbard4x4_real.txt
