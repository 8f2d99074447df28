export SCENE=cornell_box KW='{}'
ncu --set full --import-source on --clock-control none -k regex:shadeBKernel -s 2 -c 1 -o gpurun_out/s6_cornell_shadeB python tools/render_scene.py 512 512 16 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:shadeAKernel -s 2 -c 1 -o gpurun_out/s6_cornell_shadeA python tools/render_scene.py 512 512 16 1 > /dev/null 2>&1
ls -la gpurun_out/s6_cornell*
