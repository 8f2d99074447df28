# launch lists (gpu__time_duration per launch) of the three mesh scenes in the current state
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
SCENE=sphere_scene KW='{"quads":[1000,500]}' $NCU --log-file gpurun_out/${TAG}_launches_sphere1m.csv python tools/render_scene.py 1024 1024 4 1 > gpurun_out/${TAG}_sphere.log 2>&1
SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}' $NCU --log-file gpurun_out/${TAG}_launches_config3.csv python tools/render_scene.py 1920 1080 2 1 > gpurun_out/${TAG}_c3.log 2>&1
SCENE=mesh_lights_scene KW='{"num_lights":1000,"geometry_quads":[400,250],"sun":15.0,"sky":1024,"max_depth":8}' $NCU --log-file gpurun_out/${TAG}_launches_config4.csv python tools/render_scene.py 1920 1080 2 1 > gpurun_out/${TAG}_c4.log 2>&1
tail -1 gpurun_out/${TAG}_sphere.log gpurun_out/${TAG}_c3.log gpurun_out/${TAG}_c4.log
