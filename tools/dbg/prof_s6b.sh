set -x
export SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}'
# first-bounce (launch 0) and a later-bounce launch of the top and mesh kernels of config 3
ncu --set full --clock-control none --import-source on -k regex:'topKernel|meshTracePersistent' -c 6 -o gpurun_out/s6_c3_trace python tools/render_scene.py 1920 1080 1 1 > gpurun_out/s6_c3b.log 2>&1
export SCENE=sphere_scene KW='{"quads":[1000,500]}'
ncu --set full --clock-control none --import-source on -k regex:'shadeAKernel|shadeBKernel|topKernel|meshTracePersistent' -c 8 -o gpurun_out/s6_sphere python tools/render_scene.py 1024 1024 4 1 > gpurun_out/s6_sphere.log 2>&1
ls -la gpurun_out
