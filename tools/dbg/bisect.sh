python -m pytest tests/test_render_gpu.py -x -q 2>&1 | tail -3
SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}' python tools/render_scene.py 1920 1080 8 2
SCENE=cornell_box KW='{}' python tools/render_scene.py 512 512 64 3
SCENE=sphere_scene KW='{"quads":[1000,500]}' python tools/render_scene.py 1024 1024 16 3
SCENE=mesh_lights_scene KW='{"num_lights":1000,"geometry_quads":[400,250],"sun":15.0,"sky":1024,"max_depth":8}' python tools/render_scene.py 1920 1080 4 2
SCENE=sky_scene KW='{}' python tools/render_scene.py 1024 1024 16 3
