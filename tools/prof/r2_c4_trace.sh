# --set full of the fused traversal kernel on config 4: the first-bounce shadow launch (4.5 shadow rays per sample towards 1 k mesh lights) and the second closest-hit launch
export SCENE=mesh_lights_scene KW='{"num_lights":1000,"geometry_quads":[400,250],"sun":15.0,"sky":1024,"max_depth":8}'
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'sceneTracePersistent<\(bool\)1' -s 1 -c 1 -o gpurun_out/${TAG}_c4_shadow python tools/render_scene.py 1920 1080 1 1 > gpurun_out/${TAG}_c4_shadow.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'sceneTracePersistent<\(bool\)0' -s 1 -c 1 -o gpurun_out/${TAG}_c4_closest python tools/render_scene.py 1920 1080 1 1 > gpurun_out/${TAG}_c4_closest.log 2>&1
tail -n 2 gpurun_out/${TAG}_c4_shadow.log gpurun_out/${TAG}_c4_closest.log; ls -la gpurun_out/${TAG}_*
