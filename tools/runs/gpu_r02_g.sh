mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -q -rA --durations=10 > gpurun_out/r02_g_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_g_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_g_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_g_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_g_bench.json 2> gpurun_out/r02_g_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_g_bench_ref.json 2>> gpurun_out/r02_g_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r02_g_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -o gpurun_out/r02_g_pool_64spp -f python tools/profile_render.py sample 64 > gpurun_out/r02_g_ncu_pool64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -o gpurun_out/r02_g_welcome_pool_16spp -f python tools/profile_render.py welcome-2018 16 > gpurun_out/r02_g_ncu_welcome.log 2>&1
LR_BVH_TRACE=1 BUNNY_TRIS=1048576 LR_BVH_BUILDER=device python tools/profile_render.py welcome-2018 4 > gpurun_out/r02_g_bvh_trace.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitizer_probe.py > gpurun_out/r02_g_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_g_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitizer_probe.py > gpurun_out/r02_g_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_g_racecheck.log
tail -5 gpurun_out/r02_g_pytest_gpu.log; tail -2 gpurun_out/r02_g_smoke.log; tail -4 gpurun_out/r02_g_memcheck.log; tail -4 gpurun_out/r02_g_racecheck.log; cat gpurun_out/r02_g_bvh_trace.log | tail -8
