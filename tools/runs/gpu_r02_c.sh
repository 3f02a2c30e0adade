mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -q -rA --durations=10 > gpurun_out/r02_c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_c_smoke.log
LUMILLY_LIB=$PWD/lumillyrender_b200/variants/lib_bvh4.so python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q -x -k "not device_built" > gpurun_out/r02_c_pytest_bvh4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c_pytest_bvh4.log
LUMILLY_LIB=$PWD/lumillyrender_b200/variants/lib_refill12.so python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x > gpurun_out/r02_c_pytest_refill.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c_pytest_refill.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_c_bench.json 2> gpurun_out/r02_c_bench.err; echo "bench rc=$?"
SWEEP_SCENES=sample,welcome-2018 python tools/ab.py run base nostage refill12 refill8 refill96 bvh4 --rounds 2 --configs default LR_ORGANISATION=pool > gpurun_out/r02_c_ab.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -o gpurun_out/r02_c_pool_64spp -f python tools/profile_render.py sample 64 > gpurun_out/r02_c_ncu_pool64.log 2>&1
BUNNY_TRIS=1048576 python tools/profile_render.py welcome-2018 16 > gpurun_out/r02_c_l2persist.log 2>&1
BUNNY_TRIS=1048576 LR_L2_PERSIST=1 python tools/profile_render.py welcome-2018 16 >> gpurun_out/r02_c_l2persist.log 2>&1
BUNNY_TRIS=1048576 python tools/profile_render.py welcome-2018 16 >> gpurun_out/r02_c_l2persist.log 2>&1
BUNNY_TRIS=1048576 LR_L2_PERSIST=1 python tools/profile_render.py welcome-2018 16 >> gpurun_out/r02_c_l2persist.log 2>&1
tail -8 gpurun_out/r02_c_pytest_gpu.log; tail -3 gpurun_out/r02_c_pytest_bvh4.log; tail -3 gpurun_out/r02_c_pytest_refill.log; tail -2 gpurun_out/r02_c_smoke.log; cat gpurun_out/r02_c_ab.log gpurun_out/r02_c_l2persist.log
