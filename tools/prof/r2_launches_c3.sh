# launch list of config 3 under the given environment: usage TAG=.. NAME=.. r2_launches_c3.sh [VAR=val ...]
export SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}'
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_config3_${NAME}.csv python tools/render_scene.py 1920 1080 2 1 > gpurun_out/${TAG}_c3_${NAME}.log 2>&1
tail -1 gpurun_out/${TAG}_c3_${NAME}.log
