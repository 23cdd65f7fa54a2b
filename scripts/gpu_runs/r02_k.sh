#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python scripts/complex_path_check.py 4 3 6 6 > gpurun_out/r02k_complex_check.txt 2>&1; cat gpurun_out/r02k_complex_check.txt | tail -20
timeout -k 5 300 python scripts/complex_path_check.py 4 2 4 4 2>&1 | tail -16
