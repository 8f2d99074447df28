export SCENE=mesh_lights_scene KW='{"num_lights":1000,"geometry_quads":[400,250],"sun":15.0,"sky":1024,"max_depth":8}'
for L in libzyg_b200_lb8.so libzyg_b200_lb12.so libzyg_b200_lb16.so libzyg_b200_old8.so libzyg_b200_old12.so; do echo $L; ZYG_B200_LIB=$PWD/zyg_b200/$L python tools/render_scene.py 1920 1080 4 3; done
