mkdir -p gpurun_out
set -x
SWEEP_SCENES=sample,welcome-2018 python tools/ab.py run base cls16 cls24 cls32 --rounds 2 > gpurun_out/r02_h_ab.log 2>&1
LUMILLY_LIB=$PWD/lumillyrender_b200/variants/lib_cls24.so python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x > gpurun_out/r02_h_pytest_cls.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_h_pytest_cls.log
tail -3 gpurun_out/r02_h_pytest_cls.log; cat gpurun_out/r02_h_ab.log
