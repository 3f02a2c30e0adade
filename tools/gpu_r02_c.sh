mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -q -rA --durations=10 > gpurun_out/r02_c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_c_smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_c_bench.json 2> gpurun_out/r02_c_bench.err; echo "bench rc=$?"
python tools/ab.py run base nostage --rounds 2 > gpurun_out/r02_c_ab.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -o gpurun_out/r02_c_pool_64spp -f python tools/profile_render.py sample 64 > gpurun_out/r02_c_ncu_pool64.log 2>&1
BUNNY_TRIS=1048576 python tools/profile_render.py welcome-2018 16 > gpurun_out/r02_c_l2persist.log 2>&1
BUNNY_TRIS=1048576 LR_L2_PERSIST=1 python tools/profile_render.py welcome-2018 16 >> gpurun_out/r02_c_l2persist.log 2>&1
BUNNY_TRIS=1048576 python tools/profile_render.py welcome-2018 16 >> gpurun_out/r02_c_l2persist.log 2>&1
BUNNY_TRIS=1048576 LR_L2_PERSIST=1 python tools/profile_render.py welcome-2018 16 >> gpurun_out/r02_c_l2persist.log 2>&1
tail -8 gpurun_out/r02_c_pytest_gpu.log; tail -2 gpurun_out/r02_c_smoke.log; cat gpurun_out/r02_c_ab.log gpurun_out/r02_c_l2persist.log
