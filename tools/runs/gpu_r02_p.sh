mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "render_multi or resumable or aovs or accumulate or determinism" > gpurun_out/r02_p_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_p_pytest_2gpu.log
tail -12 gpurun_out/r02_p_pytest_2gpu.log
