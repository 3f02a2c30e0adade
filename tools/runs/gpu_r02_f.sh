mkdir -p gpurun_out
set -x
SWEEP_SCENES=welcome-2018,sample python tools/ab.py run base c8 c7 c5 c6s40 ggxout ggxout8 --rounds 2 > gpurun_out/r02_f_ab.log 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r02_f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_f_pytest_gpu.log
tail -4 gpurun_out/r02_f_pytest_gpu.log; cat gpurun_out/r02_f_ab.log
