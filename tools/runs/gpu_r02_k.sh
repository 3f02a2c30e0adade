mkdir -p gpurun_out
set -x
for lt in 2 3 4; do echo "== LR_LEAF_TARGET=$lt"; LR_LEAF_TARGET=$lt BVHQ_CASES=2 python tools/bvh_quality.py host 0 512; done > gpurun_out/r02_k_leaf_target.txt 2> gpurun_out/r02_k.err
cat gpurun_out/r02_k_leaf_target.txt; tail -3 gpurun_out/r02_k.err
