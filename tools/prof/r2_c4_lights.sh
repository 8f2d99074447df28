export SCENE=mesh_lights_scene KW='{"num_lights":1000,"geometry_quads":[400,250],"sun":15.0,"sky":1024,"max_depth":8}'
ncu --set full --clock-control none --import-source on -k regex:'lightSamplePersistent' -s 2 -c 2 -o gpurun_out/${TAG}_c4_lights python tools/render_scene.py 1920 1080 1 1 > gpurun_out/${TAG}_c4_lights.log 2>&1
tail -n 2 gpurun_out/${TAG}_c4_lights.log
