mkdir -p gpurun_out
set -x
SWEEP_SCENES=welcome-2018,sample python tools/ab.py run base small --rounds 3 > gpurun_out/r02_n_ab.log 2>&1
LUMILLY_LIB=$PWD/lumillyrender_b200/variants/lib_small.so python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x > gpurun_out/r02_n_pytest_small.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_n_pytest_small.log
LUMILLY_LIB=$PWD/lumillyrender_b200/variants/lib_small.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -o gpurun_out/r02_n_welcome_small_16spp -f python tools/profile_render.py welcome-2018 16 > gpurun_out/r02_n_ncu.log 2>&1
tail -3 gpurun_out/r02_n_pytest_small.log; cat gpurun_out/r02_n_ab.log
