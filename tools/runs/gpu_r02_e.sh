mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -q -rA --durations=10 > gpurun_out/r02_e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_e_pytest_gpu.log
python tools/sweep.py default LR_ORGANISATION=pool LR_ORGANISATION=persistent default > gpurun_out/r02_e_sweep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_persistent -s 1 -c 1 -o gpurun_out/r02_e_welcome_16spp -f python tools/profile_render.py welcome-2018 16 > gpurun_out/r02_e_ncu_welcome.log 2>&1
LR_ORGANISATION=pool timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -o gpurun_out/r02_e_welcome_pool_16spp -f python tools/profile_render.py welcome-2018 16 > gpurun_out/r02_e_ncu_welcome_pool.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_e_bench.json 2> gpurun_out/r02_e_bench.err; echo "bench rc=$?"
tail -6 gpurun_out/r02_e_pytest_gpu.log; cat gpurun_out/r02_e_sweep.log
