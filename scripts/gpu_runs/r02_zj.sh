#!/bin/bash
mkdir -p gpurun_out
g++ -std=c++17 -O2 -I include examples/square_fermi_hubbard.cc -L quantum_basis_b200 -lqbgpu -Wl,-rpath,$PWD/quantum_basis_b200 -o /tmp/square_fermi_hubbard || exit 3
timeout -k 5 60 /tmp/square_fermi_hubbard 4 2 4 4 > gpurun_out/r02zj_example_hubbard_4x2.txt 2>&1; echo "4x2 rc=$?"; tail -12 gpurun_out/r02zj_example_hubbard_4x2.txt
timeout -k 5 60 /tmp/square_fermi_hubbard 4 3 6 6 > gpurun_out/r02zj_example_hubbard_4x3.txt 2>&1; echo "4x3 rc=$?"; tail -4 gpurun_out/r02zj_example_hubbard_4x3.txt
