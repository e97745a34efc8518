"""A small procedural picture for the image-texture tests, and a scene that uses it on a sphere, a quad and a light."""
import numpy as np


def test_picture(w=96, h=48):
    """[h, w, 3] uint8: smooth gradients + a coarse checker + a few saturated blocks (exercises every PNG filter type)."""
    y, x = np.mgrid[0:h, 0:w]
    r = (x * 255 // max(w - 1, 1)).astype(np.uint8)
    g = (y * 255 // max(h - 1, 1)).astype(np.uint8)
    b = (((x // 8 + y // 8) % 2) * 200 + 30).astype(np.uint8)
    img = np.stack([r, g, b], axis=-1)
    img[h // 4: h // 2, w // 8: w // 4] = (255, 0, 0)
    img[h // 2: 3 * h // 4, w // 2: 5 * w // 8] = (0, 255, 64)
    return img


def textured_scene(sb, picture="pic.png", *, width=160):
    """Image-textured lambertian sphere + quad, an image-textured light, a rotated instance of a textured sphere."""
    b = sb.SceneBuilder(width=width, aspect_ratio=1.0, fov=40, center=(0, 1.5, 9), look_at=(0, 1, 0), background=(0.25, 0.3, 0.4))
    tex = b.image(picture)
    mat = b.textured(tex)
    b.place(b.sphere((-1.6, 1, 0), 1.0, mat))
    b.place(b.quad((-4, 0, -3), (8, 0, 0), (0, 0, 6), mat))
    b.place(b.quad((0.2, 0.1, -2), (2.5, 0, 0), (0, 2.5, 0), b.light(tex_idx=tex)))
    b.place(b.sphere((0, 0, 0), 0.8, mat), transform=b.transform(translation=(1.8, 0.8, 1.5), rotation_deg_axis=(40, 0, 1, 0)))
    return b
