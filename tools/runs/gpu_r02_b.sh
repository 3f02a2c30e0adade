mkdir -p gpurun_out
set -x
timeout 900 python tools/replay_divergence.py welcome-2018 1048576 2138x1536 1022,829,128,128 4 11 > gpurun_out/r02_b_divergence.log 2>&1
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_cli.py tests/test_golden.py -m gpu -q -rA --durations=10 > gpurun_out/r02_b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_b_pytest_gpu.log
SWEEP_SCENES=sample,welcome-2018 python tools/ab.py run base flatall b2 b2c5 ptd5 --rounds 2 --configs default LR_ORGANISATION=pool LR_ORGANISATION=persistent > gpurun_out/r02_b_ab.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -o gpurun_out/r02_b_pool_256spp -f python tools/profile_render.py sample 256 > gpurun_out/r02_b_ncu_pool256.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -o gpurun_out/r02_b_pool_64spp -f python tools/profile_render.py sample 64 > gpurun_out/r02_b_ncu_pool64.log 2>&1
LR_ORGANISATION=pool timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -o gpurun_out/r02_b_welcome_pool_16spp -f python tools/profile_render.py welcome-2018 16 > gpurun_out/r02_b_ncu_welcome_pool.log 2>&1
tail -3 gpurun_out/r02_b_pytest_gpu.log; cat gpurun_out/r02_b_ab.log; cat gpurun_out/r02_b_divergence.log | tail -20
