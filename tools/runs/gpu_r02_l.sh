mkdir -p gpurun_out
set -x
timeout 900 python tools/bvh_quality.py host lbvh:0 lbvh:512 ploc:0 ploc:512 > gpurun_out/r02_l_bvh_quality.txt 2> gpurun_out/r02_l_bvh_quality.err
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "device_built" > gpurun_out/r02_l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_l_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitizer_probe.py > gpurun_out/r02_l_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_l_memcheck.log
cat gpurun_out/r02_l_bvh_quality.txt; tail -3 gpurun_out/r02_l_bvh_quality.err; tail -3 gpurun_out/r02_l_pytest.log; tail -3 gpurun_out/r02_l_memcheck.log
