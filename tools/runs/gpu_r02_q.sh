mkdir -p gpurun_out
set -x
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -rA -k "flat_only" > gpurun_out/r02_q_pytest_flat.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_q_pytest_flat.log
grep "spp:" gpurun_out/r02_q_pytest_flat.log; tail -4 gpurun_out/r02_q_pytest_flat.log
