# --set full of the fused traversal kernel on config 3: the second closest-hit launch (first bounce, incoherent rays) and the second shadow launch
export SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}'
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'sceneTracePersistent<\(bool\)0' -s 10 -c 1 -o gpurun_out/${TAG}_c3_closest python tools/render_scene.py 1920 1080 4 1 > gpurun_out/${TAG}_c3_closest.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'sceneTracePersistent<\(bool\)1' -s 36 -c 1 -o gpurun_out/${TAG}_c3_shadow python tools/render_scene.py 1920 1080 4 1 > gpurun_out/${TAG}_c3_shadow.log 2>&1
tail -n 2 gpurun_out/${TAG}_c3_closest.log gpurun_out/${TAG}_c3_shadow.log; ls -la gpurun_out/${TAG}_*
