set -x
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}' $NCU --log-file gpurun_out/s6_launches_config3.csv python tools/render_scene.py 1920 1080 2 1 > gpurun_out/s6_c3.log 2>&1
SCENE=mesh_lights_scene KW='{"num_lights":1000,"geometry_quads":[400,250],"sun":15.0,"max_depth":8}' $NCU --log-file gpurun_out/s6_launches_config4.csv python tools/render_scene.py 1920 1080 1 1 > gpurun_out/s6_c4.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct -k regex:traceWide --clock-control none --csv -c 9 --log-file gpurun_out/s6_traffic_trace.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-render > gpurun_out/s6_traffic.log 2>&1
tail -2 gpurun_out/s6_c3.log gpurun_out/s6_c4.log
