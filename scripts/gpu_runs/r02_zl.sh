#!/bin/bash
# the last seconds of the round-2 GPU budget: the tests of the seven goldens added after the last full run
mkdir -p gpurun_out
timeout -k 2 14 python -m pytest tests/test_gpu_zz_step_entries.py -m gpu -q -p no:cacheprovider -k "more_reference" -rf > gpurun_out/r02zl_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02zl_pytest.log
