# Frame time of one of the bench scenes under tuning variables: usage TAG=.. r2_sweep.sh {sphere1m|config3|config4} W H SPP "VAR=1 VAR=2" ...
case "$1" in
  sphere1m) export SCENE=sphere_scene KW='{"quads":[1000,500]}';;
  config3) export SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}';;
  config4) export SCENE=mesh_lights_scene KW='{"num_lights":1000,"geometry_quads":[400,250],"sun":15.0,"sky":1024,"max_depth":8}';;
esac
name=$1; w=$2; h=$3; spp=$4; shift 4
for cfg in "$@"; do
  echo "== $name $cfg: $(env $cfg python tools/render_scene.py $w $h $spp 3 2>&1 | tail -1 | sed 's/.*spp: //')"
done | tee -a gpurun_out/${TAG}_sweep.log
