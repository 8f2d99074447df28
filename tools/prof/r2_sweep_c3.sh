# config 3 (instanced scene) frame time under tuning variables of the fused traversal kernel: usage TAG=.. r2_sweep_c3.sh "VAR=1 VAR=2" ...
export SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}'
for cfg in "$@"; do
  echo "== $cfg: $(env $cfg python tools/render_scene.py 1920 1080 4 3 2>&1 | tail -1 | sed 's/.*spp: //')"
done | tee gpurun_out/${TAG}_sweep_c3.log
