"""Procedural inputs for the BASELINE.json configs (no assets ship with the reference).

Everything is deterministic numpy; meshes are handed to the C ABI exactly the way
``src/capi-test/test.py:200-235`` feeds ``su_triangle_mesh_create``.
"""

from __future__ import annotations

import numpy as np

from .lib import RAY_DTYPE, RAY_MAX_T

# ----------------------------------------------------------------------------------------------
# value-noise displaced sphere (config 2: 1000 x 500 quads -> 1 000 000 triangles)
# ----------------------------------------------------------------------------------------------


def _hash3(ix: np.ndarray, iy: np.ndarray, iz: np.ndarray, seed: int) -> np.ndarray:
    """Integer lattice hash -> float32 in [0, 1)."""
    with np.errstate(over="ignore"):
        h = (ix.astype(np.uint32) * np.uint32(0x8DA6B343)) ^ (iy.astype(np.uint32) * np.uint32(0xD8163841)) ^ (
            iz.astype(np.uint32) * np.uint32(0xCB1AB31F)) ^ np.uint32(seed)
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x7FEB352D)
        h ^= h >> np.uint32(15)
        h *= np.uint32(0x846CA68B)
        h ^= h >> np.uint32(16)
    return (h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def _value_noise(p: np.ndarray, seed: int) -> np.ndarray:
    f = np.floor(p)
    i = f.astype(np.int64)
    t = (p - f).astype(np.float32)
    t = t * t * (np.float32(3.0) - np.float32(2.0) * t)
    ix, iy, iz = i[:, 0], i[:, 1], i[:, 2]
    out = np.zeros(p.shape[0], np.float32)
    for dx in (0, 1):
        wx = t[:, 0] if dx else (1 - t[:, 0])
        for dy in (0, 1):
            wy = t[:, 1] if dy else (1 - t[:, 1])
            for dz in (0, 1):
                wz = t[:, 2] if dz else (1 - t[:, 2])
                out += wx * wy * wz * _hash3(ix + dx, iy + dy, iz + dz, seed)
    return out


def fbm(p: np.ndarray, seed: int, octaves: int = 3) -> np.ndarray:
    amp, freq, total = 0.5, 4.0, np.zeros(p.shape[0], np.float32)
    for o in range(octaves):
        total += np.float32(amp) * (_value_noise(p * freq, seed + o) * 2 - 1)
        amp *= 0.5
        freq *= 2.0
    return total


def displaced_sphere(nu: int = 1000, nv: int = 500, radius: float = 1.0, amplitude: float = 0.05,
                     seed: int = 0x5EED0001):
    """Lat-long grid of ``nu x nv`` quads -> ``2 nu nv`` triangles, ``(nu + 1)(nv + 1)`` vertices.

    Returns ``(positions f32[n,3], normals f32[n,3], uvs f32[n,2], indices u32[m,3])``.
    """
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    uu, vv = np.meshgrid(u, v, indexing="xy")  # (nv + 1, nu + 1)
    phi = uu * 2 * np.pi
    theta = vv * np.pi
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], axis=-1).reshape(-1, 3)
    # seam: make u = 1 reuse the direction of u = 0 so the displacement is continuous
    d = d.reshape(nv + 1, nu + 1, 3)
    d[:, nu, :] = d[:, 0, :]
    d = d.reshape(-1, 3)
    r = radius + amplitude * fbm(d.astype(np.float64), seed)
    positions = (d * r[:, None]).astype(np.float32)
    normals = d.astype(np.float32)
    uvs = np.stack([uu.reshape(-1), vv.reshape(-1)], axis=-1).astype(np.float32)

    j, i = np.meshgrid(np.arange(nv), np.arange(nu), indexing="ij")
    a = (j * (nu + 1) + i).reshape(-1)
    b = a + 1
    c = a + (nu + 1)
    e = c + 1
    indices = np.empty((a.size * 2, 3), np.uint32)
    indices[0::2] = np.stack([a, c, b], axis=-1)
    indices[1::2] = np.stack([b, c, e], axis=-1)
    return positions, normals, uvs, indices


# ----------------------------------------------------------------------------------------------
# ray batches
# ----------------------------------------------------------------------------------------------


def primary_rays(width: int, height: int, eye=(0.0, 0.0, -3.0), fov_deg: float = 40.0) -> np.ndarray:
    """Pinhole rays through pixel centres, row-major, looking down +z (coherent batch)."""
    rays = np.empty(width * height, RAY_DTYPE)
    z = 0.5 * width / np.tan(0.5 * np.radians(fov_deg))
    x = (np.arange(width, dtype=np.float64) + 0.5) - 0.5 * width
    y = 0.5 * height - (np.arange(height, dtype=np.float64) + 0.5)
    xx, yy = np.meshgrid(x, y, indexing="xy")
    d = np.stack([xx, yy, np.full_like(xx, z)], axis=-1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["origin"] = np.asarray(eye, np.float32)
    rays["min_t"] = 0.0
    rays["direction"] = d.astype(np.float32)
    rays["max_t"] = RAY_MAX_T
    return rays


class PCG32:
    """Vectorised rnd.Generator (src/base/random/generator.zig:1-47): one stream per array lane."""

    MULT = np.uint64(6364136223846793005)

    def __init__(self, state: int, sequence: np.ndarray):
        seq = np.asarray(sequence, dtype=np.uint64)
        self.inc = (seq << np.uint64(1)) | np.uint64(1)
        self.state = np.zeros_like(seq)
        self.uint()
        with np.errstate(over="ignore"):
            self.state = self.state + np.uint64(state)
        self.uint()

    def uint(self) -> np.ndarray:
        old = self.state
        with np.errstate(over="ignore"):
            self.state = old * self.MULT + self.inc
        xrs = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)).astype(np.uint32)
        rot = (old >> np.uint64(59)).astype(np.uint32)
        return (xrs >> rot) | (xrs << ((np.uint32(0) - rot) & np.uint32(31)))

    def float(self) -> np.ndarray:
        bits = (self.uint() & np.uint32(0x007FFFFF)) | np.uint32(0x3F800000)
        return bits.view(np.float32) - np.float32(1.0)


def _sphere_uniform(u: np.ndarray, v: np.ndarray) -> np.ndarray:
    z = 1.0 - 2.0 * u.astype(np.float64)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = v.astype(np.float64) * (2.0 * np.pi)
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=-1)


def random_rays(n: int, radius: float = 1.5, shadow: bool = False, first: int = 0, chunk: int = 1 << 21) -> np.ndarray:
    """Incoherent batch: ray i draws from ``Generator.start(0, first + i)``.

    origin = radius * cbrt(u0) * sphere(u1, u2); closest: direction = sphere(u3, u4), max_t = RayMaxT;
    shadow: second point from (u3, u4, u5) the same way, direction = p1 - p0 (unnormalised), max_t = 1.
    """
    rays = np.empty(n, RAY_DTYPE)
    for b in range(0, n, chunk):
        e = min(n, b + chunk)
        g = PCG32(0, np.arange(first + b, first + e, dtype=np.uint64))
        u = [g.float() for _ in range(6 if shadow else 5)]
        p0 = radius * np.cbrt(u[0].astype(np.float64))[:, None] * _sphere_uniform(u[1], u[2])
        rays["origin"][b:e] = p0.astype(np.float32)
        rays["min_t"][b:e] = 0.0
        if shadow:
            p1 = radius * np.cbrt(u[3].astype(np.float64))[:, None] * _sphere_uniform(u[4], u[5])
            rays["direction"][b:e] = (p1 - p0).astype(np.float32)
            rays["max_t"][b:e] = 1.0
        else:
            rays["direction"][b:e] = _sphere_uniform(u[3], u[4]).astype(np.float32)
            rays["max_t"][b:e] = RAY_MAX_T
    return rays


# ---- config 1: Cornell box through the su_* API ------------------------------------------

def cornell_box(width=512, height=512, spp=64, max_depth=8, filter_name=None, unoccluding_light=True,
                roughness=1.0, light_value=17.0, glass=None):
    """Builds BASELINE config 1 through zyg's C API (``zyg_b200.su``): classic box x in [-1,1], y in [0,2],
    z in [-1,1], open towards -z; five Rectangle walls, one 0.5 x 0.5 Rectangle light under the ceiling, two
    rotated Cube props; camera at (0,1,-3.9) looking down +z with a 39 degree horizontal field of view.
    ``glass`` = a Glass parameter dict (e.g. ``{"ior": 1.5, "roughness": 0.0}``) turns the short box into glass and puts
    a glass ball (higher priority, different ior) half sunk into its top, so paths split, nest media and absorb.
    The engine must not be initialised yet; the caller releases it with ``su.release()``."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(39.0)))
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.0, -3.9)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth},
                                                 "light_sampling": {"split_threshold": 0.5}, "caustics": True}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})

    def substitute(color):
        return su.material_create({"rendering": {"Substitute": {"color": list(color), "roughness": roughness,
                                                                  "metallic": 0.0}}})

    white = substitute((0.73, 0.73, 0.73))
    red = substitute((0.65, 0.05, 0.05))
    green = substitute((0.12, 0.45, 0.15))
    light = su.material_create({"rendering": {"Light": {"emittance": {"spectrum": [1.0, 1.0, 1.0], "value": light_value}}}})

    walls = [
        (white, (0.0, 0.0, 0.0), (90.0, 0.0, 0.0)),    # floor, normal +y
        (white, (0.0, 2.0, 0.0), (-90.0, 0.0, 0.0)),   # ceiling, normal -y
        (white, (0.0, 1.0, 1.0), (0.0, 180.0, 0.0)),   # back wall, normal -z
        (red, (-1.0, 1.0, 0.0), (0.0, -90.0, 0.0)),    # left wall, normal +x
        (green, (1.0, 1.0, 0.0), (0.0, 90.0, 0.0)),    # right wall, normal -x
    ]
    for material, position, rotation in walls:
        prop = su.prop_create(su.RECTANGLE, [material])
        su.prop_set_transformation(prop, su.transformation(position, (2.0, 2.0, 1.0), rotation))

    tall = su.prop_create(su.CUBE, [white])
    su.prop_set_transformation(tall, su.transformation((-0.35, 0.6, 0.3), (0.6, 1.2, 0.6), (0.0, 17.0, 0.0)))
    short_material = white
    if glass is not None:
        params = {"ior": 1.5, "roughness": 0.0, "attenuation_color": [0.6, 0.9, 0.7], "attenuation_distance": 0.5}
        params.update(glass)
        short_material = su.material_create({"rendering": {"Glass": params}})
        ball_params = dict(params, ior=1.33, priority=1, attenuation_color=[0.9, 0.5, 0.4], attenuation_distance=0.25)
        if not params.get("no_ball"):
            ball_material = su.material_create({"rendering": {"Glass": ball_params}})
            ball = su.prop_create(su.SPHERE, [ball_material])
            su.prop_set_transformation(ball, su.transformation((0.35, 0.6, -0.3), (0.4, 0.4, 0.4)))
        if params.get("no_cube"):
            short_material = white
    short = su.prop_create(su.CUBE, [short_material])
    # a glass box is lifted off the floor: seen from inside, a bottom face coincident with the floor is an equal-t tie
    # that the last ulp of the ray direction decides
    su.prop_set_transformation(short, su.transformation((0.35, 0.3 if glass is None else 0.302, -0.3), (0.6, 0.6, 0.6), (0.0, -17.0, 0.0)))

    lamp = su.prop_create(su.RECTANGLE, [light], unoccluding=unoccluding_light)
    su.prop_set_transformation(lamp, su.transformation((0.0, 1.98, 0.0), (0.5, 0.5, 1.0), (-90.0, 0.0, 0.0)))
    su.light_create(lamp)
    return camera


def sphere_scene(width=512, height=512, spp=16, max_depth=8, filter_name=None, quads=(1000, 500), metallic=0.0,
                 roughness=0.6, sun=None):
    """The config-2 mesh (displaced sphere, 2 * quads[0] * quads[1] triangles) as a path-traced scene: the mesh sits
    on a diffuse floor inside a large open room lit by one Rectangle light. Returns the number of meshes."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(40.0)))
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 0.0, -3.0)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth}}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})

    grey = su.material_create({"rendering": {"Substitute": {"color": [0.6, 0.6, 0.6], "roughness": 1.0}}})
    body = su.material_create({"rendering": {"Substitute": {"color": [0.8, 0.45, 0.2], "roughness": roughness,
                                                            "metallic": metallic}}})
    light = su.material_create({"rendering": {"Light": {"emittance": {"value": 12.0}}}})

    positions, normals, uvs, indices = displaced_sphere(*quads)
    # displaced_sphere winds its triangles clockwise seen from outside; a one-sided Substitute needs the geometric
    # normal (cross(b - a, c - a), triangle_data.zig:140-149) on the side of the shading normals
    indices = np.ascontiguousarray(indices.reshape(-1, 3)[:, [0, 2, 1]])
    shape = su.triangle_mesh_create(positions, indices, normals, uvs)
    su.prop_create(shape, [body])

    floor = su.prop_create(su.RECTANGLE, [grey])
    su.prop_set_transformation(floor, su.transformation((0.0, -1.1, 0.0), (12.0, 12.0, 1.0), (90.0, 0.0, 0.0)))
    wall = su.prop_create(su.RECTANGLE, [grey])
    su.prop_set_transformation(wall, su.transformation((0.0, 2.0, 3.0), (12.0, 8.0, 1.0), (0.0, 180.0, 0.0)))

    lamp = su.prop_create(su.RECTANGLE, [light], unoccluding=True)
    su.prop_set_transformation(lamp, su.transformation((1.5, 3.0, -1.0), (1.5, 1.5, 1.0), (-90.0, 0.0, 0.0)))
    su.light_create(lamp)
    if sun is not None:
        # a Distant light (shape id 3): the prop's -z axis points at the sun, scale.x = tan of its angular radius
        sun_material = su.material_create({"rendering": {"Light": {"emittance": {"spectrum": [1.0, 0.9, 0.75], "value": float(sun)}}}})
        sun_prop = su.prop_create(su.DISTANT, [sun_material])
        su.prop_set_transformation(sun_prop, su.transformation((0.0, 0.0, 0.0), (0.05, 0.05, 0.05), (-55.0, -35.0, 0.0)))
        su.light_create(sun_prop)
    return 1


def many_lights_scene(width=256, height=256, spp=16, max_depth=6, num_lights=64, seed=5, filter_name=None,
                      split_threshold=0.5, unoccluding=True):
    """A closed room (6 x 3 x 6) with `num_lights` small Rectangle lights of random position, orientation, size, colour
    and power (a quarter of them two-sided), a few diffuse and glossy cubes: the light tree has many nodes, the adaptive
    split returns several picks near the lights and one far away."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(70.0)))
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.4, -2.8)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth},
                                                 "light_sampling": {"split_threshold": split_threshold}}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})

    wall = su.material_create({"rendering": {"Substitute": {"color": [0.7, 0.7, 0.7], "roughness": 1.0}}})
    glossy = su.material_create({"rendering": {"Substitute": {"color": [0.8, 0.6, 0.3], "roughness": 0.35, "metallic": 1.0}}})
    for position, scale, rotation in [((0, 0, 0), (6, 6, 1), (90, 0, 0)), ((0, 3, 0), (6, 6, 1), (-90, 0, 0)),
                                      ((0, 1.5, 3), (6, 3, 1), (0, 180, 0)), ((0, 1.5, -3), (6, 3, 1), (0, 0, 0)),
                                      ((-3, 1.5, 0), (6, 3, 1), (0, -90, 0)), ((3, 1.5, 0), (6, 3, 1), (0, 90, 0))]:
        prop = su.prop_create(su.RECTANGLE, [wall])
        su.prop_set_transformation(prop, su.transformation(tuple(map(float, position)), tuple(map(float, scale)), tuple(map(float, rotation))))
    for k, (x, z) in enumerate([(-1.2, 0.8), (0.2, 1.6), (1.4, 0.2)]):
        cube = su.prop_create(su.CUBE, [glossy if 1 == k else wall])
        su.prop_set_transformation(cube, su.transformation((x, 0.4, z), (0.8, 0.8, 0.8), (0.0, 25.0 * k, 0.0)))

    rng = PCG32(0, np.array([seed], np.uint64))
    for i in range(num_lights):
        r = [float(rng.float()[0]) for _ in range(10)]
        colour = [0.3 + 0.7 * r[0], 0.3 + 0.7 * r[1], 0.3 + 0.7 * r[2]]
        params = {"emittance": {"spectrum": colour, "value": 3.0 + 27.0 * r[3] * r[3]}}
        if 0 == i % 4:
            params["two_sided"] = True
        material = su.material_create({"rendering": {"Light": params}})
        lamp = su.prop_create(su.RECTANGLE, [material], unoccluding=unoccluding)
        size = 0.1 + 0.3 * r[4]
        su.prop_set_transformation(lamp, su.transformation((-2.7 + 5.4 * r[5], 0.3 + 2.5 * r[6], -2.7 + 5.4 * r[7]), (size, size, 1.0),
                                                           (-90.0 + 120.0 * (r[8] - 0.5), 360.0 * r[9], 0.0)))
        su.light_create(lamp)
    return camera


def sphere_lights_scene(width=128, height=128, spp=16, max_depth=6, filter_name=None, split_threshold=0.5, num_samples=1,
                        unoccluding=(True, False, True)):
    """A closed room lit by three Sphere lights (Sphere.sampleTo / pdf / emission, sphere.zig:271-279, 323-393, 472-487): a
    large one close to the floor, a small one and a tiny far one (the small-angle branch of the cone sampling); each either
    un-occluding (a scene file's default for Light entities: gathered by Prop.emission) or an ordinary occluding prop."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(70.0)))
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.4, -2.8)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth},
                                                 "light_sampling": {"split_threshold": split_threshold}}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})

    wall = su.material_create({"rendering": {"Substitute": {"color": [0.7, 0.7, 0.7], "roughness": 1.0}}})
    glossy = su.material_create({"rendering": {"Substitute": {"color": [0.8, 0.6, 0.3], "roughness": 0.35, "metallic": 1.0}}})
    walls = [((0, 0, 0), (6, 6, 1), (90, 0, 0)), ((0, 3, 0), (6, 6, 1), (-90, 0, 0)),
             ((0, 1.5, 3), (6, 3, 1), (0, 180, 0)), ((0, 1.5, -3), (6, 3, 1), (0, 0, 0)),
             ((-3, 1.5, 0), (6, 3, 1), (0, -90, 0)), ((3, 1.5, 0), (6, 3, 1), (0, 90, 0))]
    for position, scale, rotation in walls:
        prop = su.prop_create(su.RECTANGLE, [wall])
        su.prop_set_transformation(prop, su.transformation(tuple(map(float, position)), tuple(map(float, scale)), tuple(map(float, rotation))))
    for k, (x, z) in enumerate([(-1.2, 0.8), (1.4, 0.2)]):
        cube = su.prop_create(su.CUBE, [glossy if 1 == k else wall])
        su.prop_set_transformation(cube, su.transformation((x, 0.4, z), (0.8, 0.8, 0.8), (0.0, 25.0 * k, 0.0)))
    lamps = [((0.3, 0.9, 1.4), 1.0, [1.0, 0.8, 0.6], 3.0), ((-1.6, 2.2, -0.5), 0.25, [0.6, 0.8, 1.0], 60.0),
             ((2.4, 2.7, 2.5), 0.02, [1.0, 1.0, 1.0], 6000.0)]
    for (position, diameter, colour, value), un in zip(lamps, unoccluding):
        material = su.material_create({"rendering": {"Light": {"emittance": {"spectrum": colour, "value": value, "num_samples": num_samples}}}})
        lamp = su.prop_create(su.SPHERE, [material], unoccluding=un)
        su.prop_set_transformation(lamp, su.transformation(position, (diameter, diameter, diameter)))
        su.light_create(lamp)
    return 0


def checker_image(size=64, cells=8, a=(0.8, 0.8, 0.8), b=(0.15, 0.3, 0.6), dtype=np.float32):
    """A checkerboard colour texture: float32 RGB (ACEScg as is) or uint8 (read as sRGB like Texture.Byte3_sRGB)."""
    i = np.arange(size) * cells // size
    mask = ((i[:, None] + i[None, :]) & 1).astype(bool)
    img = np.where(mask[..., None], np.asarray(a, np.float64), np.asarray(b, np.float64))
    return (img * 255.0 + 0.5).astype(np.uint8) if dtype == np.uint8 else np.ascontiguousarray(img, np.float32)


def textured_scene(width=128, height=128, spp=16, max_depth=6, filter_name=None, quads=(48, 24), uniform=None, nearest=False):
    """Substitute colour maps (substitute_material.zig:120, texture_sampler.zig:63-170): a float checker on the ground
    (Repeat addressing, texture scale 4), an sRGB byte checker on a cube, a float checker on a displaced-sphere mesh through
    its own uvs. `uniform` = (r, g, b) replaces every image by that constant colour (then the film must equal the one of plain
    uniform-colour materials). Returns the number of meshes."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(55.0)))
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.6, -4.2), rotation_deg=(-14.0, 0.0, 0.0)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth}}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})

    sampler = {"filter": "Nearest"} if nearest else {}

    def textured(image, roughness, metallic=0.0, scale=1.0, address="Repeat"):
        if uniform is not None:
            return su.material_create({"rendering": {"Substitute": {"color": {"sRGB": list(uniform)} if image.dtype == np.uint8 else list(uniform),
                                                                    "roughness": roughness, "metallic": metallic}}})
        iid = su.image_create(image)
        return su.material_create({"rendering": {"Substitute": {
            "color": {"id": iid, "scale": scale, "sampler": dict(sampler, address=address)}, "roughness": roughness, "metallic": metallic}}})

    ground = textured(checker_image(64, 8), 1.0, scale=4.0)
    g = su.prop_create(su.RECTANGLE, [ground])
    su.prop_set_transformation(g, su.transformation((0.0, 0.0, 0.0), (12.0, 12.0, 1.0), (90.0, 0.0, 0.0)))
    crate = textured(checker_image(32, 4, (0.9, 0.6, 0.2), (0.3, 0.1, 0.05), np.uint8), 0.6, address="Clamp")
    cube = su.prop_create(su.CUBE, [crate])
    su.prop_set_transformation(cube, su.transformation((-1.3, 0.5, 0.4), (1.0, 1.0, 1.0), (0.0, 30.0, 0.0)))
    ball_material = textured(checker_image(128, 16, (0.9, 0.9, 0.9), (0.7, 0.1, 0.1)), 0.35, metallic=0.0)
    positions, normals, uvs, indices = displaced_sphere(*quads, seed=0x5EED0009)
    indices = np.ascontiguousarray(indices.reshape(-1, 3)[:, [0, 2, 1]])
    ball = su.prop_create(su.triangle_mesh_create(positions, indices, normals, uvs), [ball_material])
    su.prop_set_transformation(ball, su.transformation((1.2, 0.8, 0.0), (0.75, 0.75, 0.75), (0.0, 20.0, 0.0)))

    light = su.material_create({"rendering": {"Light": {"emittance": {"value": 25.0}}}})
    lamp = su.prop_create(su.RECTANGLE, [light], unoccluding=True)
    su.prop_set_transformation(lamp, su.transformation((0.0, 4.0, -0.5), (2.0, 2.0, 1.0), (-90.0, 0.0, 0.0)))
    su.light_create(lamp)
    return 1


def bump_normal_map(size=64, waves=4.0, strength=0.5, dtype=np.float32):
    """Tangent-space xy of the normals of a sinusoidal height field, (size, size, 2): float32, or uint8 snorm (enc.floatToSnorm8)."""
    t = (np.arange(size) + 0.5) / size * (2.0 * np.pi * waves)
    nx = -strength * np.cos(t)[None, :] * np.ones((size, 1))
    ny = -strength * np.cos(t)[:, None] * np.ones((1, size))
    xy = np.stack([nx, ny], -1)
    xy /= np.sqrt(1.0 + (xy * xy).sum(-1, keepdims=True))
    if dtype == np.uint8:
        return ((xy + 1.0) * np.where(xy > 0.0, 127.5, 128.0)).astype(np.uint8)
    return np.ascontiguousarray(xy, np.float32)


def surface_maps_scene(width=128, height=128, spp=16, max_depth=6, filter_name=None, quads=(48, 24), uniform=False, nearest=False,
                       maps=("roughness", "metallic", "normal")):
    """Substitute roughness, metallic and normal maps (substitute_material.zig:122-123, 157-159; material_helper.zig:16-79): a ground
    plane with a float roughness ramp and a byte (snorm) normal map, a cube with byte (unorm) roughness and metallic checkers, a
    displaced-sphere mesh with a float normal map over its own uvs and a metallic map. `uniform` replaces every map by the constant it
    averages to (a flat normal map is no map). Returns the number of meshes."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(55.0)))
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.6, -4.2), rotation_deg=(-14.0, 0.0, 0.0)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth}}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})
    sampler = {"filter": "Nearest"} if nearest else {}

    def mapped(kind, image, constant, scale=1.0, address="Repeat"):
        if uniform or kind not in maps:
            return constant
        return {"id": su.image_create(image), "scale": scale, "sampler": dict(sampler, address=address)}

    def material(color, roughness, metallic, normal=None):
        desc = {"color": color, "roughness": roughness, "metallic": metallic}
        if isinstance(normal, dict):
            desc["normal"] = normal
        return su.material_create({"rendering": {"Substitute": desc}})

    ramp = np.ascontiguousarray(np.tile(np.linspace(0.08, 0.9, 64, dtype=np.float32)[None, :], (64, 1)))
    ground = material([0.6, 0.6, 0.55], mapped("roughness", ramp, 0.49, scale=3.0), 0.0,
                      mapped("normal", bump_normal_map(64, 4.0, 0.6, np.uint8), None, scale=6.0))
    g = su.prop_create(su.RECTANGLE, [ground])
    su.prop_set_transformation(g, su.transformation((0.0, 0.0, 0.0), (12.0, 12.0, 1.0), (90.0, 0.0, 0.0)))

    i = np.arange(32) * 4 // 32
    checker = ((i[:, None] + i[None, :]) & 1).astype(np.uint8)
    crate = material([0.9, 0.6, 0.2], mapped("roughness", (60 + 150 * checker).astype(np.uint8), 0.53, address="Clamp"),
                     mapped("metallic", (255 * checker).astype(np.uint8), 0.5, address="Clamp"))
    cube = su.prop_create(su.CUBE, [crate])
    su.prop_set_transformation(cube, su.transformation((-1.3, 0.5, 0.4), (1.0, 1.0, 1.0), (0.0, 30.0, 0.0)))

    j = np.arange(128) * 16 // 128
    spots = (((j[:, None] + j[None, :]) & 1) * 1.0).astype(np.float32)
    ball_material = material([0.8, 0.75, 0.7], 0.3, mapped("metallic", spots, 0.5),
                             mapped("normal", bump_normal_map(128, 12.0, 0.8), None))
    positions, normals, uvs, indices = displaced_sphere(*quads, seed=0x5EED0009)
    indices = np.ascontiguousarray(indices.reshape(-1, 3)[:, [0, 2, 1]])
    ball = su.prop_create(su.triangle_mesh_create(positions, indices, normals, uvs), [ball_material])
    su.prop_set_transformation(ball, su.transformation((1.2, 0.8, 0.0), (0.75, 0.75, 0.75), (0.0, 20.0, 0.0)))

    light = su.material_create({"rendering": {"Light": {"emittance": {"value": 25.0}}}})
    lamp = su.prop_create(su.RECTANGLE, [light], unoccluding=True)
    su.prop_set_transformation(lamp, su.transformation((0.0, 4.0, -0.5), (2.0, 2.0, 1.0), (-90.0, 0.0, 0.0)))
    su.light_create(lamp)
    return 1


def coated_scene(width=128, height=128, spp=16, max_depth=6, filter_name=None, quads=(48, 24), coated=True, thickness=0.02, normal_map=True):
    """Clear-coated Substitutes (substitute_coating.zig, substitute_sample.zig:138-142, 304-336, 412-433): a glossy ground with a smooth
    colourless coat, a diffuse cube under a thick amber coat (absorption along both passes through the layer), a rough-metal Sphere with
    a rough coat, a mesh whose base has a normal map while the coat keeps the interpolated normal. `coated=False` renders the same
    bases without their coats. Returns the number of meshes."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(55.0)))
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.6, -4.2), rotation_deg=(-14.0, 0.0, 0.0)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth}}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})

    def material(base, coating):
        desc = dict(base)
        if coated:
            desc["coating"] = coating
        return su.material_create({"rendering": {"Substitute": desc}})

    ground = material({"color": [0.5, 0.05, 0.04], "roughness": 0.5, "metallic": 0.0},
                      {"thickness": thickness, "roughness": 0.02, "ior": 1.5})
    g = su.prop_create(su.RECTANGLE, [ground])
    su.prop_set_transformation(g, su.transformation((0.0, 0.0, 0.0), (12.0, 12.0, 1.0), (90.0, 0.0, 0.0)))

    amber = material({"color": [0.8, 0.8, 0.8], "roughness": 1.0, "metallic": 0.0},
                     {"thickness": 10.0 * thickness, "roughness": 0.1, "ior": 1.6, "color": [0.9, 0.6, 0.15], "attenuation_distance": 0.1})
    cube = su.prop_create(su.CUBE, [amber])
    su.prop_set_transformation(cube, su.transformation((-1.3, 0.5, 0.4), (1.0, 1.0, 1.0), (0.0, 30.0, 0.0)))

    metal = material({"color": [0.95, 0.64, 0.54], "roughness": 0.35, "metallic": 1.0}, {"thickness": thickness, "roughness": 0.3, "ior": 1.4})
    ball = su.prop_create(su.SPHERE, [metal])
    su.prop_set_transformation(ball, su.transformation((-0.1, 0.45, -1.2), (0.45, 0.45, 0.45)))

    base = {"color": [0.1, 0.25, 0.6], "roughness": 0.4, "metallic": 0.0, "anisotropy": 0.5}
    if normal_map:
        base["normal"] = {"id": su.image_create(bump_normal_map(128, 12.0, 0.8))}
    paint = material(base, {"thickness": thickness, "roughness": 0.05, "ior": 1.5})
    positions, normals, uvs, indices = displaced_sphere(*quads, seed=0x5EED0009)
    indices = np.ascontiguousarray(indices.reshape(-1, 3)[:, [0, 2, 1]])
    mesh = su.prop_create(su.triangle_mesh_create(positions, indices, normals, uvs), [paint])
    su.prop_set_transformation(mesh, su.transformation((1.2, 0.8, 0.0), (0.75, 0.75, 0.75), (0.0, 20.0, 0.0)))

    light = su.material_create({"rendering": {"Light": {"emittance": {"value": 25.0}}}})
    lamp = su.prop_create(su.RECTANGLE, [light], unoccluding=True)
    su.prop_set_transformation(lamp, su.transformation((0.0, 4.0, -0.5), (2.0, 2.0, 1.0), (-90.0, 0.0, 0.0)))
    su.light_create(lamp)
    return 1


def disk_scene(width=128, height=128, spp=16, max_depth=6, filter_name=None, disk_lights=False, split_threshold=0.5, num_samples=1):
    """Disk props (disk.zig:28-134): a glossy disk table top on a diffuse floor, a tilted metal disk that casts a round shadow, and an
    emissive disk that is not registered as a light (its emission is met by the paths, without next-event estimation), under a
    Rectangle lamp. With `disk_lights` the lamp is an un-occluding Disk (Disk.emission, disk.zig:171-179) and the small emissive disk is
    registered as a light too (Disk.sampleTo with equi-angular sampling and Disk.pdf, disk.zig:181-332, 492-533)."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(55.0)))
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.6, -4.2), rotation_deg=(-14.0, 0.0, 0.0)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth},
                                                 "light_sampling": {"split_threshold": split_threshold}}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})

    floor = su.material_create({"rendering": {"Substitute": {"color": [0.55, 0.55, 0.5], "roughness": 1.0}}})
    g = su.prop_create(su.RECTANGLE, [floor])
    su.prop_set_transformation(g, su.transformation((0.0, 0.0, 0.0), (12.0, 12.0, 1.0), (90.0, 0.0, 0.0)))

    top = su.material_create({"rendering": {"Substitute": {"color": [0.2, 0.35, 0.6], "roughness": 0.25, "two_sided": True}}})
    table = su.prop_create(su.DISK, [top])
    su.prop_set_transformation(table, su.transformation((-0.9, 0.6, 0.3), (2.0, 2.0, 1.0), (90.0, 0.0, 0.0)))

    metal = su.material_create({"rendering": {"Substitute": {"color": [0.95, 0.8, 0.5], "roughness": 0.3, "metallic": 1.0, "two_sided": True}}})
    mirror = su.prop_create(su.DISK, [metal])
    su.prop_set_transformation(mirror, su.transformation((1.3, 1.1, 0.8), (1.6, 1.6, 1.0), (20.0, -35.0, 0.0)))

    glow = su.material_create({"rendering": {"Substitute": {"color": [0.0, 0.0, 0.0], "roughness": 1.0, "two_sided": True,
                                                             "emittance": {"value": 3.0}}}})
    lamp_disk = su.prop_create(su.DISK, [glow])
    su.prop_set_transformation(lamp_disk, su.transformation((0.4, 0.35, -0.9), (0.7, 0.7, 1.0), (0.0, 0.0, 0.0)))
    if disk_lights:
        su.light_create(lamp_disk)

    light = su.material_create({"rendering": {"Light": {"emittance": {"value": 25.0, "num_samples": num_samples}}}})
    lamp = su.prop_create(su.DISK if disk_lights else su.RECTANGLE, [light], unoccluding=True)
    su.prop_set_transformation(lamp, su.transformation((0.0, 4.0, -0.5), (2.0, 2.0, 1.0), (-90.0, 0.0, 0.0)))
    su.light_create(lamp)
    return 0


def image_light_scene(width=128, height=128, spp=16, max_depth=5, filter_name=None, split_threshold=0.5, num_samples=1,
                      image=None, value=6.0, two_sided=False, unoccluding=True):
    """A closed room lit by a Rectangle whose Light material carries an emission image (a PropImage light on a finite shape:
    Rectangle.sampleMaterialTo / materialPdf, texels picked through the material's Distribution2D) — a "stained-glass" ceiling
    panel. `image` = None uses a colour checker; a constant image makes it an ordinary area light sampled another way."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(70.0)))
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.4, -2.8)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth},
                                                 "light_sampling": {"split_threshold": split_threshold}}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})

    wall = su.material_create({"rendering": {"Substitute": {"color": [0.7, 0.7, 0.7], "roughness": 1.0}}})
    glossy = su.material_create({"rendering": {"Substitute": {"color": [0.8, 0.6, 0.3], "roughness": 0.35, "metallic": 1.0}}})
    walls = [((0, 0, 0), (6, 6, 1), (90, 0, 0)), ((0, 3, 0), (6, 6, 1), (-90, 0, 0)),
             ((0, 1.5, 3), (6, 3, 1), (0, 180, 0)), ((0, 1.5, -3), (6, 3, 1), (0, 0, 0)),
             ((-3, 1.5, 0), (6, 3, 1), (0, -90, 0)), ((3, 1.5, 0), (6, 3, 1), (0, 90, 0))]
    for position, scale, rotation in walls:
        prop = su.prop_create(su.RECTANGLE, [wall])
        su.prop_set_transformation(prop, su.transformation(tuple(map(float, position)), tuple(map(float, scale)), tuple(map(float, rotation))))
    for k, (x, z) in enumerate([(-1.2, 0.8), (1.4, 0.2)]):
        cube = su.prop_create(su.CUBE, [glossy if 1 == k else wall])
        su.prop_set_transformation(cube, su.transformation((x, 0.4, z), (0.8, 0.8, 0.8), (0.0, 25.0 * k, 0.0)))

    if image is None:
        image = checker_image(32, 4, (1.0, 0.2, 0.1), (0.1, 0.3, 1.0))
    emittance = {"value": float(value), "num_samples": num_samples}
    if image is not False:  # False: a plain uniform Rectangle light (spherical-rectangle sampling) for comparison
        emittance["emission_map"] = {"id": su.image_create(image), "sampler": {"address": "Clamp"}}
    material = su.material_create({"rendering": {"Light": {"two_sided": two_sided, "emittance": emittance}}})
    panel = su.prop_create(su.RECTANGLE, [material], unoccluding=unoccluding)
    su.prop_set_transformation(panel, su.transformation((0.2, 2.6, 0.6), (2.2, 1.4, 1.0), (-90.0, 0.0, 20.0)))
    su.light_create(panel)
    return 0


def icosahedron():
    """Unit icosahedron: (positions f32[12,3], indices u32[20,3]), counter-clockwise seen from outside."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], np.uint32)
    return v.astype(np.float32), f


def procedural_sky(size=1024, sun_rotation_deg=(-70.0, -20.0, 0.0), zenith=(0.25, 0.45, 1.0), horizon=(0.9, 0.95, 1.0), glow=6.0):
    """A sky radiance image in the Canopy's mapping (equidistant hemisphere -> unit disk -> [0, 1]^2, canopy.zig:164-202), the
    layout zyg bakes its own sky model into (sky.zig:44,176-343; the arpraguesky dataset is not shipped): zenith-to-horizon
    gradient plus a glow around the sun direction, black outside the disk. Returns (size, size, 3) float32 (ACEScg)."""
    c = (np.arange(size, dtype=np.float64) + 0.5) / size
    u, v = np.meshgrid(c, c)
    dx, dy = 2.0 * u - 1.0, 2.0 * v - 1.0
    r = np.sqrt(dx * dx + dy * dy)
    inside = r <= 1.0
    colat = np.minimum(r, 1.0) * (np.pi / 2.0)
    lon = np.arctan2(-dy, dx)
    d = np.stack([np.sin(colat) * np.cos(lon), np.sin(colat) * np.sin(lon), np.cos(colat)], -1)  # canopy-local, z = up
    # the sun points along -r[2] of its Distant prop (distant.zig:22-54); world -> canopy-local with the sky rotation (90, 0, 0)
    from . import su

    sun_dir = -su.transformation(rotation_deg=sun_rotation_deg)[2, :3].astype(np.float64)
    sky_rot = su.transformation(rotation_deg=(90.0, 0.0, 0.0))[:3, :3].astype(np.float64)
    sun_local = sky_rot @ sun_dir
    t = np.cos(colat)[..., None] ** 0.6
    img = (1.0 - t) * np.asarray(horizon) + t * np.asarray(zenith)
    cos_sun = np.clip(d @ sun_local, -1.0, 1.0)
    img = img * (0.35 + glow * np.exp((cos_sun - 1.0) * 40.0)[..., None])
    img[~inside] = 0.0
    return np.ascontiguousarray(img, np.float32)


def add_sky(image: np.ndarray, value: float = 1.0):
    """Sky dome the way sky.zig:57-66,121-131 sets it up, with a Light material carrying the image as its emission map
    (the C API has no Sky material): Canopy prop rotated 90 degrees about X (z = up), clamped texture addressing."""
    from . import su

    sky_image = su.image_create(image)
    material = su.material_create({"rendering": {"Light": {"emittance": {
        "emission_map": {"id": sky_image, "sampler": {"address": "Clamp"}}, "value": float(value)}}}})
    sky = su.prop_create(su.CANOPY, [material])
    su.prop_set_transformation(sky, su.transformation(rotation_deg=(90.0, 0.0, 0.0)))
    su.light_create(sky)
    return sky


def sky_scene(width=256, height=256, spp=16, max_depth=6, filter_name=None, sky_size=256, sun=8.0, sky_value=1.0, uniform_sky=None,
              ground_albedo=0.5, objects=True, split_threshold=0.5):
    """Open scene lit by an image-mapped Canopy sky (a PropImage light, importance-sampled through Distribution2D) and
    optionally a Distant sun: two infinite lights, so Tree.infinite_light_distribution is exercised too. `uniform_sky` = L
    replaces the image by a constant radiance L (furnace-style checks: a ground of albedo a shows a * L)."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(60.0)))
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.2, -4.0), rotation_deg=(-10.0, 0.0, 0.0)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth}, "light_sampling": {"split_threshold": split_threshold}}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})

    ground = su.material_create({"rendering": {"Substitute": {"color": [ground_albedo] * 3, "roughness": 1.0}}})
    g = su.prop_create(su.RECTANGLE, [ground])
    su.prop_set_transformation(g, su.transformation((0.0, 0.0, 0.0), (400.0, 400.0, 1.0), (90.0, 0.0, 0.0)))
    if objects:
        red = su.material_create({"rendering": {"Substitute": {"color": [0.7, 0.2, 0.15], "roughness": 1.0}}})
        gold = su.material_create({"rendering": {"Substitute": {"color": [0.9, 0.7, 0.3], "roughness": 0.3, "metallic": 1.0}}})
        cube = su.prop_create(su.CUBE, [red])
        su.prop_set_transformation(cube, su.transformation((-1.0, 0.5, 0.5), (1.0, 1.0, 1.0), (0.0, 30.0, 0.0)))
        ball = su.prop_create(su.SPHERE, [gold])
        su.prop_set_transformation(ball, su.transformation((1.1, 0.6, 0.0), (1.2, 1.2, 1.2)))
    if uniform_sky is not None:
        image = np.full((sky_size, sky_size, 3), float(uniform_sky), np.float32)
    else:
        image = procedural_sky(sky_size)
    add_sky(image, sky_value)
    if sun is not None:
        sun_material = su.material_create({"rendering": {"Light": {"emittance": {"spectrum": [1.0, 0.9, 0.75], "value": float(sun) * 400.0}}}})
        sun_prop = su.prop_create(su.DISTANT, [sun_material])
        su.prop_set_transformation(sun_prop, su.transformation((0.0, 0.0, 0.0), (0.05, 0.05, 0.05), (-70.0, -20.0, 0.0)))
        su.light_create(sun_prop)
    return 0


def mesh_lights_scene(width=256, height=256, spp=16, max_depth=6, num_lights=24, seed=11, filter_name=None,
                      split_threshold=0.5, big_light_quads=(24, 12), geometry_quads=None, sun=None, unoccluding=False, sky=None):
    """Config-4 style: a closed room lit only by emissive triangle meshes - `num_lights` small icosahedra (20 triangles,
    radius 0.05-0.12, instances of one mesh with PCG-random colour and power) and one larger emissive displaced sphere
    with many triangles, so both the scene light tree and the per-part primitive trees (spherical-triangle sampling near,
    area sampling far) are exercised. `geometry_quads` adds a diffuse displaced sphere of 2 * nu * nv triangles (config 4: 200 k
    triangles of diffuse geometry), `sun` a Distant light. Returns the number of meshes."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(70.0)))
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.4, -2.8)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth},
                                                 "light_sampling": {"split_threshold": split_threshold}}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})

    wall = su.material_create({"rendering": {"Substitute": {"color": [0.7, 0.7, 0.7], "roughness": 1.0}}})
    glossy = su.material_create({"rendering": {"Substitute": {"color": [0.8, 0.6, 0.3], "roughness": 0.35, "metallic": 1.0}}})
    walls = [((0, 0, 0), (6, 6, 1), (90, 0, 0)), ((0, 3, 0), (6, 6, 1), (-90, 0, 0)),
             ((0, 1.5, 3), (6, 3, 1), (0, 180, 0)), ((0, 1.5, -3), (6, 3, 1), (0, 0, 0)),
             ((-3, 1.5, 0), (6, 3, 1), (0, -90, 0)), ((3, 1.5, 0), (6, 3, 1), (0, 90, 0))]
    if sun is not None:
        del walls[1]  # no ceiling: the sun (and escaping paths) get in and out
    for position, scale, rotation in walls:
        prop = su.prop_create(su.RECTANGLE, [wall])
        su.prop_set_transformation(prop, su.transformation(tuple(map(float, position)), tuple(map(float, scale)), tuple(map(float, rotation))))
    for k, (x, z) in enumerate([(-1.2, 0.8), (1.4, 0.2)]):
        cube = su.prop_create(su.CUBE, [glossy if 1 == k else wall])
        su.prop_set_transformation(cube, su.transformation((x, 0.4, z), (0.8, 0.8, 0.8), (0.0, 25.0 * k, 0.0)))

    positions, indices = icosahedron()
    ico = su.triangle_mesh_create(positions, indices, positions.copy(), np.zeros((positions.shape[0], 2), np.float32))
    rng = PCG32(0, np.array([seed], np.uint64))
    for i in range(num_lights):
        r = [float(rng.float()[0]) for _ in range(8)]
        colour = [0.3 + 0.7 * r[0], 0.3 + 0.7 * r[1], 0.3 + 0.7 * r[2]]
        material = su.material_create({"rendering": {"Light": {"emittance": {"spectrum": colour, "value": (10.0 + 60.0 * r[3] * r[3]) * min(1.0, 24.0 / num_lights)}}}})
        lamp = su.prop_create(ico, [material], unoccluding=unoccluding)  # scene files make Light entities un-occluding by default
        radius = 0.05 + 0.07 * r[4]
        su.prop_set_transformation(lamp, su.transformation((-2.6 + 5.2 * r[5], 0.4 + 2.3 * r[6], -2.6 + 5.2 * r[7]), (radius, radius, radius),
                                                           (0.0, 360.0 * r[0], 0.0)))
        su.light_create(lamp)

    bp, bn, buv, bi = displaced_sphere(*big_light_quads, seed=0x5EED0042)
    bi = np.ascontiguousarray(bi.reshape(-1, 3)[:, [0, 2, 1]])
    big = su.triangle_mesh_create(bp, bi, bn, buv)
    big_material = su.material_create({"rendering": {"Light": {"emittance": {"spectrum": [1.0, 0.85, 0.6], "value": 4.0}}}})
    big_prop = su.prop_create(big, [big_material])
    su.prop_set_transformation(big_prop, su.transformation((0.2, 0.45, 1.5), (0.45, 0.45, 0.45), (0.0, 30.0, 0.0)))
    su.light_create(big_prop)
    num_meshes = 2
    if geometry_quads is not None:
        gp, gn, guv, gi = displaced_sphere(*geometry_quads, seed=0x5EED0077, amplitude=0.12)
        gi = np.ascontiguousarray(gi.reshape(-1, 3)[:, [0, 2, 1]])
        geometry = su.triangle_mesh_create(gp, gi, gn, guv)
        geometry_prop = su.prop_create(geometry, [wall])
        su.prop_set_transformation(geometry_prop, su.transformation((-0.6, 0.9, 1.2), (0.9, 0.9, 0.9), (0.0, 0.0, 0.0)))
        num_meshes += 1
    if sun is not None:
        sun_material = su.material_create({"rendering": {"Light": {"emittance": {"spectrum": [1.0, 0.9, 0.75], "value": float(sun)}}}})
        sun_prop = su.prop_create(su.DISTANT, [sun_material])
        su.prop_set_transformation(sun_prop, su.transformation((0.0, 0.0, 0.0), (0.05, 0.05, 0.05), (-70.0, -20.0, 0.0)))
        su.light_create(sun_prop)
    if sky is not None:  # config 4: the sky dome (baked 1024^2 image on a Canopy) next to the sun
        add_sky(procedural_sky(int(sky)), 1.0)
    return num_meshes


def instanced_scene(width=512, height=512, spp=16, max_depth=8, filter_name=None, grid=(32, 32), prototypes=4,
                    quads=(100, 50), seed=3, glass=True, sun=None, instancer=None):
    """Config-3 style scene through the C API: `prototypes` displaced-sphere meshes (seeds 1..), instanced
    grid[0] x grid[1] times with su_prop_create_instance on a jittered grid (uniform scale 0.3-0.6, random Y rotation,
    PCG32 stream `seed`), a ground Rectangle, one Rectangle light and optionally a Distant sun. Materials go by prototype
    id mod 3 like BASELINE config 3: diffuse Substitute, gold-like rough metal, Glass (ior 1.5, smooth, attenuation
    distance 1; a glossy dielectric Substitute when `glass` is off). Returns the number of meshes."""
    from . import su

    su.init()
    camera = su.perspective_camera_create(width, height)
    su.camera_set_fov(float(np.radians(50.0)))
    extent = 0.5 * max(grid)
    su.prop_set_transformation(camera, su.transformation(position=(0.0, 0.45 * extent, -1.15 * extent),
                                                         rotation_deg=(-28.0, 0.0, 0.0)))
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": max_depth}}}})
    su.sensor_create({"filter": {filter_name: {}}} if filter_name else {})

    ground = su.material_create({"rendering": {"Substitute": {"color": [0.55, 0.55, 0.5], "roughness": 1.0}}})
    palette = [
        {"Substitute": {"color": [0.7, 0.25, 0.2], "roughness": 1.0, "metallic": 0.0}},
        {"Substitute": {"color": [1.0, 0.77, 0.34], "roughness": 0.3, "metallic": 1.0}},
        {"Glass": {"ior": 1.5, "roughness": 0.0, "attenuation_color": [0.75, 0.9, 0.95], "attenuation_distance": 1.0}} if glass
        else {"Substitute": {"color": [0.2, 0.5, 0.75], "roughness": 0.15, "metallic": 0.0}},
    ]
    materials = [su.material_create({"rendering": palette[i % len(palette)]}) for i in range(prototypes)]
    light = su.material_create({"rendering": {"Light": {"emittance": {"value": 40.0}}}})

    protos = []
    for i in range(prototypes):
        positions, normals, uvs, indices = displaced_sphere(*quads, seed=0x5EED0001 + i)
        indices = np.ascontiguousarray(indices.reshape(-1, 3)[:, [0, 2, 1]])
        shape = su.triangle_mesh_create(positions, indices, normals, uvs)
        protos.append(su.prop_create(shape, [materials[i]]))

    rng = PCG32(0, np.array([seed], np.uint64))
    indices, matrices = [], []
    for gy in range(grid[1]):
        for gx in range(grid[0]):
            r = [float(rng.float()[0]) for _ in range(5)]
            index = int(r[0] * prototypes) % prototypes
            scale = 0.3 + 0.3 * r[1]
            x = (gx + 0.5 + 0.6 * (r[2] - 0.5)) - 0.5 * grid[0]
            z = (gy + 0.5 + 0.6 * (r[3] - 0.5)) - 0.5 * grid[1]
            matrix = su.transformation((x, 1.05 * scale, z), (scale, scale, scale), (0.0, 360.0 * r[4], 0.0))
            if instancer is None:
                su.prop_set_transformation(su.prop_create_instance(protos[index]), matrix)
            else:
                indices.append(index)
                matrices.append(matrix)
    if instancer is None:
        # the prototypes themselves stay out of the picture (they are props too: park them below the ground, invisible)
        for proto in protos:
            su.prop_set_transformation(proto, su.transformation((0.0, -50.0, 0.0), (0.01, 0.01, 0.01)))
            su.prop_set_visibility(proto, False, False)
    else:
        # a scene file's "Instancer" entity (scene_loader.zig:401-508): prototypes + instance transformations relative to the
        # instancer, `instancer` = the instancer's own transformation (a 4x4 as su.transformation returns it)
        entity = su.instancer_create(protos, indices, np.stack(matrices))
        su.prop_set_transformation(entity, instancer)

    floor = su.prop_create(su.RECTANGLE, [ground])
    su.prop_set_transformation(floor, su.transformation((0.0, 0.0, 0.0), (3.0 * max(grid), 3.0 * max(grid), 1.0), (90.0, 0.0, 0.0)))
    lamp = su.prop_create(su.RECTANGLE, [light], unoccluding=True)
    su.prop_set_transformation(lamp, su.transformation((0.25 * extent, 1.2 * extent, 0.0), (0.6 * extent, 0.6 * extent, 1.0), (-90.0, 0.0, 0.0)))
    su.light_create(lamp)
    if sun is not None:
        sun_material = su.material_create({"rendering": {"Light": {"emittance": {"spectrum": [1.0, 0.9, 0.75], "value": float(sun)}}}})
        sun_prop = su.prop_create(su.DISTANT, [sun_material])
        su.prop_set_transformation(sun_prop, su.transformation((0.0, 0.0, 0.0), (0.05, 0.05, 0.05), (-55.0, -35.0, 0.0)))
        su.light_create(sun_prop)
    return prototypes
