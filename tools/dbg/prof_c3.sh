export SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}'
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s6_launches_config3_b.csv python tools/render_scene.py 1920 1080 2 1 > gpurun_out/s6_c3c.log 2>&1
python tools/render_scene.py 1920 1080 8 3
