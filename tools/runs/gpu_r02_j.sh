mkdir -p gpurun_out
set -x
python tools/bvh_quality.py > gpurun_out/r02_j_bvh_quality.txt 2> gpurun_out/r02_j_bvh_quality.err
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "device_built" > gpurun_out/r02_j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_j_pytest.log
cat gpurun_out/r02_j_bvh_quality.txt; tail -3 gpurun_out/r02_j_bvh_quality.err; tail -3 gpurun_out/r02_j_pytest.log
