mkdir -p gpurun_out
set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_persistent -s 1 -c 1 -o gpurun_out/r02_r_brdf_64spp -f python tools/profile_render.py brdf 64 > gpurun_out/r02_r_ncu_brdf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_persistent -s 1 -c 1 -o gpurun_out/r02_r_newcbox_64spp -f python tools/profile_render.py new-cbox 64 > gpurun_out/r02_r_ncu_newcbox.log 2>&1
tail -2 gpurun_out/r02_r_ncu_brdf.log gpurun_out/r02_r_ncu_newcbox.log
