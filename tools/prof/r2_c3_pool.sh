# --set full of the ray-pool variant of the fused kernel on config 3 (second closest-hit launch of a 4-spp pass)
export SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}' ZYGPU_SCENE_POOL=1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'scenePoolTrace<\(bool\)0' -s 10 -c 1 -o gpurun_out/${TAG}_c3_pool_closest python tools/render_scene.py 1920 1080 4 1 > gpurun_out/${TAG}_c3_pool_closest.log 2>&1
tail -n 2 gpurun_out/${TAG}_c3_pool_closest.log
