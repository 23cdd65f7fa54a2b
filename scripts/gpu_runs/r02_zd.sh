#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_resume.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r02zd_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r02zd_pytest.log
