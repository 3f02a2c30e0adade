mkdir -p gpurun_out
set -x
python tools/full_film_parity.py 64 > gpurun_out/r02_i_full_film_parity.txt 2> gpurun_out/r02_i_full_film_parity.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_i_bench_driver_like.json 2> gpurun_out/r02_i_bench.err ) 2> gpurun_out/r02_i_bench_time.txt
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_i_bench_ref_driver_like.json 2>> gpurun_out/r02_i_bench.err ) 2>> gpurun_out/r02_i_bench_time.txt
cat gpurun_out/r02_i_full_film_parity.txt gpurun_out/r02_i_bench_time.txt; tail -3 gpurun_out/r02_i_full_film_parity.err
