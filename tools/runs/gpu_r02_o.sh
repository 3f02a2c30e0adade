mkdir -p gpurun_out
set -x
SWEEP_SCENES=welcome-2018 python tools/ab.py run base divin allin coldin --rounds 3 > gpurun_out/r02_o_ab.log 2>&1
cat gpurun_out/r02_o_ab.log
