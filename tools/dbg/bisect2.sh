export SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}'
(cd _bisect/ef5839c && ncu --set full --import-source on --clock-control none -k regex:shadeAKernel -c 1 -o ../../gpurun_out/s6_shadeA5_old python tools/render_scene.py 1920 1080 1 1 > /dev/null 2>&1)
ncu --set full --import-source on --clock-control none -k regex:shadeAKernel -c 1 -o gpurun_out/s6_shadeA5_new python tools/render_scene.py 1920 1080 1 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
