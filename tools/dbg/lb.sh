for L in libzyg_b200.so libzyg_b200_tb6.so libzyg_b200_tb8.so libzyg_b200_tb10.so; do echo $L
SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}' ZYG_B200_LIB=$PWD/zyg_b200/$L python tools/render_scene.py 1920 1080 4 2
SCENE=mesh_lights_scene KW='{"num_lights":1000,"geometry_quads":[400,250],"sun":15.0,"sky":1024,"max_depth":8}' ZYG_B200_LIB=$PWD/zyg_b200/$L python tools/render_scene.py 1920 1080 4 2
SCENE=cornell_box ZYG_B200_LIB=$PWD/zyg_b200/$L python tools/render_scene.py 512 512 64 3
done
