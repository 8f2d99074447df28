export SCENE=mesh_lights_scene KW='{"num_lights":1000,"geometry_quads":[400,250],"sun":15.0,"sky":1024,"max_depth":8}'
ncu --set full --clock-control none --import-source on -k regex:'lightSamplePersistent|lightSelectPersistent' -c 4 -o gpurun_out/s6_c4_lights python tools/render_scene.py 1920 1080 1 1 > gpurun_out/s6_c4b.log 2>&1
tail -2 gpurun_out/s6_c4b.log
python tools/render_scene.py 1920 1080 4 2
