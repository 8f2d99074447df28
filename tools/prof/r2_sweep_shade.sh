# shade kernels compiled for 4 / 5 / 6 resident blocks per SM (libraries built with -DZYGPU_SHADE_BLOCKS=n next to the default one)
for lib in "" _sb5 _sb6; do
  for scene in "cornell_box 512 512 64 {}" "sphere_scene 1024 1024 16 {\"quads\":[1000,500]}"; do
    set -- $scene
    grid=4; [ "$lib" = "_sb5" ] && grid=5; [ "$lib" = "_sb6" ] && grid=6
    echo "== lib$lib $1: $(ZYG_B200_LIB=$PWD/zyg_b200/libzyg_b200$lib.so ZYGPU_SHADE_GRID=$grid SCENE=$1 KW="$5" python tools/render_scene.py $2 $3 $4 3 2>&1 | tail -n 1 | sed 's/.*spp: //')"
  done
done | tee gpurun_out/${TAG}_sweep_shade.log
