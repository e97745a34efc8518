"""Scene authoring for the repo's JSON scene format (loader: src/Serialize.cpp:199-360).

Counterpart of the reference's scene script (`make_scene.py`): it builds the same kind of documents —
textures / materials / primitives / a scene graph of nodes with optional transforms — and the standard test scenes
(Cornell box, Cornell box with constant media, the book-2 final scene, a random sphere field).  Everything here is
host-side convenience: the output is a plain dict / JSON string that `Scene.from_builder` / `Scene.from_string` (rt2_scene_load_string) compiles.

    b = SceneBuilder(width=600, aspect_ratio=1.0, fov=40, center=(278, 278, -800), look_at=(278, 278, 0))
    white = b.lambertian((0.73, 0.73, 0.73))
    b.place(b.quad((0, 0, 0), (555, 0, 0), (0, 0, 555), white))
    scene = raytrace2_b200.Scene.from_builder(b)
"""
from __future__ import annotations

import json
import math
import random
from typing import Iterable, Optional, Sequence

Vec = Sequence[float]


def _v(x: Vec) -> list:
    return [float(c) if not float(c).is_integer() else int(c) for c in x]


class SceneBuilder:
    """Accumulates the five arrays of a scene file plus the camera block.  Every add-method returns the index the JSON
    refers to (materials by `material`, textures by `tex_idx`, primitives by a node's `primitive`)."""

    def __init__(self, *, width: int = 600, aspect_ratio: float = 1.0, fov: int = 40, center: Vec = (0, 0, 1), look_at: Vec = (0, 0, 0),
                 defocus_angle: float = 0.0, focus_distance: float = 10.0, background: Vec = (0, 0, 0)):
        self.camera = {"fov": int(fov), "center": _v(center), "look_at": _v(look_at), "width": int(width), "aspect_ratio": float(aspect_ratio)}
        if defocus_angle:
            self.camera["defocus_angle"] = float(defocus_angle)
            self.camera["focus_distance"] = float(focus_distance)
        self.background = _v(background)
        self.textures: list = []
        self.materials: list = []
        self.primitives: list = []
        self.nodes: list = []

    # ---- textures (Serialize.cpp:216-242) ----
    def solid(self, albedo: Vec) -> int:
        return self._push(self.textures, {"type": "solid_color", "albedo": _v(albedo)})

    def checker(self, scale: float, even_tex: int, odd_tex: int) -> int:
        return self._push(self.textures, {"type": "checker", "scale": float(scale), "even_tex_idx": int(even_tex), "odd_tex_idx": int(odd_tex)})

    def noise(self, scale: float, *, marble: bool = True, albedo: Vec = (1, 1, 1), point_count: int = 256) -> int:
        return self._push(self.textures, {"type": "noise", "scale": float(scale), "noise_type": 1 if marble else 0, "albedo": _v(albedo),
                                          "point_count": int(point_count)})

    def image(self, path: str) -> int:
        """Image texture (schema extension, include/rt2.h RT2_TEX_IMAGE): PNG / PPM file, path relative to the data directory."""
        return self._push(self.textures, {"type": "image", "path": str(path)})

    # ---- materials (Serialize.cpp:244-285) ----
    def lambertian(self, albedo: Vec) -> int:
        return self._push(self.materials, {"type": "lambertian", "albedo": _v(albedo)})

    def metal(self, albedo: Vec, fuzz: float = 0.0) -> int:
        return self._push(self.materials, {"type": "metal", "albedo": _v(albedo), "fuzz": float(fuzz)})

    def dielectric(self, refraction_index: float) -> int:
        return self._push(self.materials, {"type": "dielectric", "refraction_index": float(refraction_index)})

    def textured(self, tex_idx: int) -> int:
        return self._push(self.materials, {"type": "texture", "tex_idx": int(tex_idx)})

    def light(self, albedo: Optional[Vec] = None, *, tex_idx: Optional[int] = None) -> int:
        m = {"type": "diffuse_light"}
        if tex_idx is not None:
            m["tex_idx"] = int(tex_idx)
        else:
            m["albedo"] = _v(albedo if albedo is not None else (1, 1, 1))
        return self._push(self.materials, m)

    # ---- primitives (Serialize.cpp:287-342) ----
    def sphere(self, center: Vec, radius: float, material: int, *, displacement: Optional[Vec] = None, medium: Optional[dict] = None) -> int:
        p = {"type": "sphere", "center": _v(center), "radius": float(radius), "material": int(material)}
        if displacement is not None:
            p["displacement"] = _v(displacement)
        return self._push(self.primitives, self._with_medium(p, medium))

    def quad(self, q: Vec, u: Vec, v: Vec, material: int, *, medium: Optional[dict] = None) -> int:
        return self._push(self.primitives, self._with_medium({"type": "quad", "q": _v(q), "u": _v(u), "v": _v(v), "material": int(material)}, medium))

    def box(self, a: Vec, b: Vec, material: int, *, medium: Optional[dict] = None) -> int:
        return self._push(self.primitives, self._with_medium({"type": "box", "a": _v(a), "b": _v(b), "material": int(material)}, medium))

    @staticmethod
    def constant_medium(density: float, albedo: Vec) -> dict:
        """The optional `constant_medium` block of a primitive: the primitive becomes the boundary (ConstantMedium.cpp:10-12)."""
        return {"density": float(density), "albedo": _v(albedo)}

    # ---- scene graph (Serialize.cpp:161-197) ----
    @staticmethod
    def transform(*, translation: Optional[Vec] = None, rotation_deg_axis: Optional[Vec] = None, scale: Optional[Vec] = None) -> dict:
        t = {}
        if translation is not None:
            t["translation"] = _v(translation)
        if rotation_deg_axis is not None:
            t["rotation"] = _v(rotation_deg_axis)  # [degrees, axis x, y, z]
        if scale is not None:
            t["scale"] = _v(scale)
        return t

    @staticmethod
    def node(primitive: Optional[int] = None, *, transform: Optional[dict] = None, children: Optional[Iterable[dict]] = None) -> dict:
        n = {}
        if primitive is not None:
            n["primitive"] = int(primitive)
        if transform:
            n["transform"] = transform
        if children:
            n["children"] = list(children)
        return n

    def place(self, primitive: Optional[int] = None, *, transform: Optional[dict] = None, children: Optional[Iterable[dict]] = None) -> dict:
        """Appends a top-level node (an entry of the `scene` array) and returns it."""
        n = self.node(primitive, transform=transform, children=children)
        self.nodes.append(n)
        return n

    # ---- output ----
    def to_dict(self) -> dict:
        return {"textures": self.textures, "materials": self.materials, "primitives": self.primitives, "scene": self.nodes,
                "camera": self.camera, "background_color": self.background}

    def to_json(self, **kw) -> str:
        return json.dumps(self.to_dict(), **kw)

    def write(self, path: str) -> None:
        with open(path, "w") as f:
            f.write(self.to_json())

    @staticmethod
    def _with_medium(p: dict, medium: Optional[dict]) -> dict:
        if medium is not None:
            p["constant_medium"] = medium
        return p

    @staticmethod
    def _push(arr: list, item: dict) -> int:
        arr.append(item)
        return len(arr) - 1


def write_settings(path: str, *, num_samples: int, max_depth: int = 50, headless: bool = True) -> None:
    """The AppSettings file the app reads (Serialize.cpp:56-65); headless = render num_samples, save a PNG, exit."""
    with open(path, "w") as f:
        json.dump({"num_samples": int(num_samples), "max_depth": int(max_depth), "render_once": bool(headless),
                   "save_after_render_once": bool(headless), "render_window": not headless}, f)


# ---------------------------------------------------------------------------------------------------------------------
# standard scenes
# ---------------------------------------------------------------------------------------------------------------------
def _cornell_shell(b: SceneBuilder):
    """Walls and light of the repo's Cornell scenes (data/cornell_original_test.json): 555-unit room, 330 x 305 light of
    emission 7 just below the ceiling."""
    red, white, green = b.lambertian((0.65, 0.05, 0.05)), b.lambertian((0.73, 0.73, 0.73)), b.lambertian((0.12, 0.45, 0.15))
    lamp = b.light((7, 7, 7))
    walls = [
        b.quad((555, 0, 0), (0, 555, 0), (0, 0, 555), green),
        b.quad((0, 0, 0), (0, 555, 0), (0, 0, 555), red),
        b.quad((113, 554, 127), (330, 0, 0), (0, 0, 305), lamp),
        b.quad((0, 0, 0), (555, 0, 0), (0, 0, 555), white),
        b.quad((0, 555, 0), (555, 0, 0), (0, 0, 555), white),
        b.quad((0, 0, 555), (555, 0, 0), (0, 555, 0), white),
    ]
    return red, white, walls


def cornell_box(*, width: int = 600) -> SceneBuilder:
    """The Cornell box of BASELINE config 1: five walls, a light quad, two boxes as rotated + translated instances."""
    b = SceneBuilder(width=width, aspect_ratio=1.0, fov=40, center=(278, 278, -800), look_at=(278, 278, 0), background=(0, 0, 0))
    _, _, walls = _cornell_shell(b)
    for w in walls:
        b.place(w)
    box_white = b.lambertian((0.73, 0.73, 0.73))
    cube = b.box((0, 0, 0), (165, 165, 165), box_white)
    tall = b.box((0, 0, 0), (165, 330, 165), box_white)
    b.place(cube, transform=b.transform(translation=(130, 0, 65), rotation_deg_axis=(15, 0, 1, 0)))
    b.place(tall, transform=b.transform(translation=(265, 0, 295), rotation_deg_axis=(-18, 0, 1, 0)))
    return b


def cornell_volume(*, width: int = 600, density: float = 0.01) -> SceneBuilder:
    """BASELINE config 3: the two boxes are boundaries of constant media (white smoke and black smoke)."""
    b = SceneBuilder(width=width, aspect_ratio=1.0, fov=40, center=(278, 278, -800), look_at=(278, 278, 0), background=(0, 0, 0))
    red, _, walls = _cornell_shell(b)
    cube = b.box((0, 0, 0), (165, 165, 165), red, medium=b.constant_medium(density, (1, 1, 1)))
    tall = b.box((0, 0, 0), (165, 330, 165), red, medium=b.constant_medium(density, (0, 0, 0)))
    b.place(cube, transform=b.transform(translation=(130, 0, 65), rotation_deg_axis=(-18, 0, 1, 0)))
    b.place(tall, transform=b.transform(translation=(265, 0, 295), rotation_deg_axis=(15, 0, 1, 0)))
    for w in walls:
        b.place(w)
    return b


def book2_final(*, width: int = 600, seed: int = 7, boxes_per_side: int = 20, cluster_spheres: int = 1000) -> SceneBuilder:
    """The 'final scene' of the second book in the repo's format (BASELINE config 4): a grid of random-height boxes, an area
    light, a moving sphere, glass and metal spheres, a sphere-bounded blue medium inside glass, thin fog around everything,
    a marble sphere, and a rotated + translated cluster of small spheres held by one node (one instance)."""
    rng = random.Random(seed)
    b = SceneBuilder(width=width, aspect_ratio=1.0, fov=40, center=(478, 278, -600), look_at=(278, 278, 0), background=(0, 0, 0))
    ground = b.lambertian((0.48, 0.83, 0.53))
    w = 100.0
    for i in range(boxes_per_side):
        for j in range(boxes_per_side):
            x0, z0 = -1000.0 + i * w, -1000.0 + j * w
            b.place(b.box((x0, 0, z0), (x0 + w, rng.uniform(1, 101), z0 + w), ground))
    b.place(b.quad((123, 554, 147), (300, 0, 0), (0, 0, 265), b.light((7, 7, 7))))
    b.place(b.sphere((400, 400, 200), 50, b.lambertian((0.7, 0.3, 0.1)), displacement=(30, 0, 0)))
    glass = b.dielectric(1.5)
    b.place(b.sphere((260, 150, 45), 50, glass))
    b.place(b.sphere((0, 150, 145), 50, b.metal((0.8, 0.8, 0.9), 1.0)))
    b.place(b.sphere((360, 150, 145), 70, glass))
    b.place(b.sphere((360, 150, 145), 70, glass, medium=b.constant_medium(0.2, (0.2, 0.4, 0.9))))
    b.place(b.sphere((0, 0, 0), 5000, glass, medium=b.constant_medium(0.0001, (1, 1, 1))))
    b.place(b.sphere((220, 280, 300), 80, b.textured(b.noise(0.2, marble=True))))
    white = b.lambertian((0.73, 0.73, 0.73))
    cluster = [b.node(b.sphere((rng.uniform(0, 165), rng.uniform(0, 165), rng.uniform(0, 165)), 10, white)) for _ in range(cluster_spheres)]
    b.place(transform=b.transform(translation=(-100, 270, 395), rotation_deg_axis=(15, 0, 1, 0)), children=cluster)
    return b


def random_spheres(n: int = 480, *, width: int = 1200, aspect_ratio: float = 16 / 9, seed: int = 11) -> SceneBuilder:
    """A book-1 style field of small random spheres on a huge ground sphere (lambertian / metal / glass), with depth of field."""
    rng = random.Random(seed)
    b = SceneBuilder(width=width, aspect_ratio=aspect_ratio, fov=20, center=(13, 2, 3), look_at=(0, 0, 0), defocus_angle=0.6, focus_distance=10.0,
                     background=(0.7, 0.8, 1.0))
    b.place(b.sphere((0, -1000, 0), 1000, b.lambertian((0.5, 0.5, 0.5))))
    side = max(1, int(math.ceil(math.sqrt(n))))
    placed = 0
    for a in range(-side // 2, side - side // 2):
        for c in range(-side // 2, side - side // 2):
            if placed >= n:
                break
            centre = (a + 0.9 * rng.random(), 0.2, c + 0.9 * rng.random())
            pick = rng.random()
            if pick < 0.8:
                mat = b.lambertian((rng.random() * rng.random(), rng.random() * rng.random(), rng.random() * rng.random()))
            elif pick < 0.95:
                mat = b.metal((rng.uniform(0.5, 1), rng.uniform(0.5, 1), rng.uniform(0.5, 1)), rng.uniform(0, 0.5))
            else:
                mat = b.dielectric(1.5)
            b.place(b.sphere(centre, 0.2, mat))
            placed += 1
    b.place(b.sphere((0, 1, 0), 1.0, b.dielectric(1.5)))
    b.place(b.sphere((-4, 1, 0), 1.0, b.lambertian((0.4, 0.2, 0.1))))
    b.place(b.sphere((4, 1, 0), 1.0, b.metal((0.7, 0.6, 0.5), 0.0)))
    return b
