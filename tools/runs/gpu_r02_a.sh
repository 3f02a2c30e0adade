mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_a_bench.json 2> gpurun_out/r02_a_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_a_bench_ref.json 2>> gpurun_out/r02_a_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r02_a_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -o gpurun_out/r02_a_pool_1024spp -f python tools/profile_render.py sample 1024 > gpurun_out/r02_a_ncu_pool.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_persistent -s 1 -c 1 -o gpurun_out/r02_a_welcome_64spp -f python tools/profile_render.py welcome-2018 64 > gpurun_out/r02_a_ncu_welcome.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_smoke.log | tail -3
