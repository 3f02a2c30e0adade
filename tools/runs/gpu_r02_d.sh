mkdir -p gpurun_out
set -x
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "render_multi" > gpurun_out/r02_d_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_d_pytest_multi.log
LR_MULTI_TRACE=1 timeout 900 python tools/multi_probe.py 1024 8 > gpurun_out/r02_d_multi_probe.log 2> gpurun_out/r02_d_multi_trace.log
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_d_bench_n$n.json 2> gpurun_out/r02_d_bench_n$n.err
done
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02_d_bench_n1.json 2> gpurun_out/r02_d_bench_n1.err
tail -5 gpurun_out/r02_d_pytest_multi.log; cat gpurun_out/r02_d_multi_probe.log; for n in 1 2 4 8; do python -c "
import json,sys
l=json.load(open('gpurun_out/r02_d_bench_n$n.json'))
print($n, l['value'], l['e2e']['value'], l['ms_per_step'], l['scaling'])
"; done
