for L in libzyg_b200.so libzyg_b200_fi.so; do echo "== $L"
SCENE=mesh_lights_scene KW='{"num_lights":1000,"geometry_quads":[400,250],"sun":15.0,"sky":1024,"max_depth":8}' ZYG_B200_LIB=$PWD/zyg_b200/$L python tools/render_scene.py 1920 1080 4 2
SCENE=mesh_lights_scene KW='{}' ZYG_B200_LIB=$PWD/zyg_b200/$L python tools/render_scene.py 1024 1024 16 2
done
ZYG_B200_LIB=$PWD/zyg_b200/libzyg_b200_fi.so python -m pytest tests/test_render_gpu.py -x -q -k "mesh_lights or many_lights" 2>&1 | tail -2
