mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -q -rA --durations=8 > gpurun_out/r02_m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_m_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_m_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_m_smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_m_bench.json 2> gpurun_out/r02_m_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_m_pytest_gpu.log; tail -2 gpurun_out/r02_m_smoke.log; grep "device build" gpurun_out/r02_m_pytest_gpu.log
