python -m pytest tests/test_render_gpu.py -x -q 2>&1 | tail -2
SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}' python tools/render_scene.py 1920 1080 8 2
SCENE=cornell_box KW='{}' python tools/render_scene.py 512 512 64 3
SCENE=cornell_box KW='{}' python tools/render_scene.py 512 512 8 3
SCENE=sphere_scene KW='{"quads":[1000,500]}' python tools/render_scene.py 1024 1024 4 3
