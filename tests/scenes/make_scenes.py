#!/usr/bin/env python3
"""Generate the synthetic test / benchmark scenes (JSON in the reference's scene schema + OBJ meshes).

The schema is the one the reference's loader reads (Demo/SceneLoader.cpp:692-820: materials / objects / lights /
camera; shapes 418-499; lights 501-618).  Everything is generated programmatically and deterministically
(`random.seed(1)`), nothing is copied from the reference's Data/ directory.

    python tests/scenes/make_scenes.py <out_dir> [--sponza]    # --sponza also writes the 262k-triangle bench mesh

Scenes:
    c1_sphere_area   BASELINE config 1: diffuse sphere + rect area light
    cornell          BASELINE config 2: Cornell box (boxes, 2 spheres, rect area light) - same geometry values as
                     the reference's Data/TestScenes/cornell_box.json so the CPU baseline is comparable
    caustics         BASELINE config 5 in miniature: rough-glass sphere + glass block under a small area light (VCM tests)
    materials        all nine BSDFs on spheres/boxes, area(sphere) + point + spot + background lights
    mesh_small       ~2.3k-triangle OBJ (two materials from a .mtl) + analytic metal/glass spheres, background +
                     directional light (the two-light pattern of Data/TestScenes/sponza.json:19-31)
    textures         every texture kind of the reference's loader (checkerboard, noise, bitmap, nested mix; bitmap files
                     referenced by name from materials; 24-bit, 8-bit palette and 8-bit grey BMPs generated here) on
                     base colour / emission / roughness / metalness / normal map, a textured OBJ (vt + map_Kd + norm),
                     a rect with textureScale, an environment-mapped background light, a missing texture file
    sponza_class     BASELINE config 3/4: ~262k-triangle atrium (floor/ceiling height fields, two rows of faceted
                     columns, arches, drapes), diffuse mesh materials + metal + dielectric analytic spheres,
                     background + 1-degree directional light
"""
import json
import math
import os
import random
import sys


def mat(name, bsdf, base, **kw):
    m = {"name": name, "bsdf": bsdf, "baseColor": [float(c) for c in base]}
    m.update(kw)
    return m


def xform(t=None, o=None):
    d = {}
    if t is not None:
        d["translation"] = [float(v) for v in t]
    if o is not None:
        d["orientation"] = [float(v) for v in o]
    return d


def scene_c1():
    return {
        "materials": [mat("white", "diffuse", (0.8, 0.8, 0.8))],
        "objects": [{"type": "sphere", "radius": 1.0, "transform": xform((0, 0, 0)), "material": "white"}],
        "lights": [{"type": "area", "color": [10.0, 10.0, 10.0],
                    "transform": xform((0.0, 3.0, 0.0), (90.0, 0.0, 0.0)),
                    "shape": {"type": "plane", "size": [1.0, 1.0]}}],
        "camera": {"transform": xform((0.0, 1.0, -5.0), (10.0, 0.0, 0.0)), "fieldOfView": 45.0},
    }


def scene_cornell():
    box = lambda size, t, m, o=None: {"type": "box", "size": [float(s) for s in size], "transform": xform(t, o), "material": m}
    return {
        "materials": [
            mat("white", "diffuse", (0.8, 0.8, 0.8)), mat("red", "diffuse", (0.8, 0.1, 0.1)),
            mat("green", "diffuse", (0.1, 0.2, 0.8)), mat("glass", "dielectric", (1.0, 1.0, 1.0)),
            mat("gold", "metal", (1.0, 0.6, 0.1)), mat("silver", "metal", (0.98, 0.98, 0.98)),
            mat("glossy", "roughMetal", (1.0, 1.0, 1.0), roughness=0.3),
        ],
        "objects": [
            box((5.0, 0.5, 5.0), (0.0, 5.0, 0.0), "white"), box((5.0, 0.5, 5.0), (0.0, -5.0, 0.0), "white"),
            box((0.5, 5.0, 5.0), (5.0, 0.0, 0.0), "green"), box((0.5, 5.0, 5.0), (-5.0, 0.0, 0.0), "red"),
            box((5.0, 5.0, 0.5), (0.0, 0.0, -5.0), "white"),
            box((1.5, 2.0, 1.5), (-2.2, -3.5, -0.8), "white", (0.0, 0.74, 0.0)),
            box((1.5, 2.0, 1.5), (2.0, -3.5, 1.0), "white", (0.0, 1.35, 0.0)),
            {"type": "sphere", "radius": 1.5, "transform": xform((-2.2, 0.002, -0.8)), "material": "glass"},
            {"type": "sphere", "radius": 1.5, "transform": xform((2.0, 0.002, 1.0)), "material": "silver"},
        ],
        "lights": [{"type": "area", "color": [5.0, 5.0, 5.0],
                    "transform": xform((0.0, 4.45, 0.0), (90.0, 0.0, 0.0)),
                    "shape": {"type": "plane", "size": [2.0, 2.0]}}],
        "camera": {"transform": xform((-0.1, 0.2, 12.0), (0.01, 180.0, 0.0)), "fieldOfView": 55.0},
    }


def scene_materials():
    mats = [
        mat("floor", "diffuse", (0.7, 0.7, 0.7)),
        mat("m_diffuse", "diffuse", (0.8, 0.3, 0.2)),
        mat("m_roughDiffuse", "roughDiffuse", (0.2, 0.7, 0.3), roughness=0.6),
        mat("m_dielectric", "dielectric", (0.9, 0.95, 1.0), IoR=1.45),
        mat("m_roughDielectric", "roughDielectric", (0.95, 0.9, 0.8), roughness=0.25, IoR=1.6),
        mat("m_metal", "metal", (0.95, 0.7, 0.3), IoR=0.2, K=3.5),
        mat("m_roughMetal", "roughMetal", (0.9, 0.9, 0.95), roughness=0.35),
        mat("m_plastic", "plastic", (0.1, 0.3, 0.9), IoR=1.5),
        mat("m_roughPlastic", "roughPlastic", (0.9, 0.2, 0.6), roughness=0.4, IoR=1.5),
        mat("m_emissive", "null", (0.0, 0.0, 0.0), emissionColor=[2.0, 1.0, 0.5]),
        mat("m_smoothRough", "roughMetal", (0.8, 0.8, 0.8), roughness=0.001),
    ]
    objs = [{"type": "box", "size": [12.0, 0.25, 8.0], "transform": xform((0.0, -1.25, 0.0)), "material": "floor"},
            {"type": "box", "size": [12.0, 4.0, 0.25], "transform": xform((0.0, 2.5, 4.0)), "material": "floor"}]
    names = ["m_diffuse", "m_roughDiffuse", "m_dielectric", "m_roughDielectric", "m_metal", "m_roughMetal",
             "m_plastic", "m_roughPlastic", "m_emissive", "m_smoothRough"]
    for i, n in enumerate(names):
        x = -6.75 + 1.5 * i
        if i % 3 == 2:
            objs.append({"type": "box", "size": [0.5, 0.6, 0.5], "transform": xform((x, -0.4, 0.3 * (i % 2)), (0.0, 25.0 + 7.0 * i, 0.0)), "material": n})
        else:
            objs.append({"type": "sphere", "radius": 0.6, "transform": xform((x, -0.4, -0.2 * (i % 2))), "material": n})
    objs.append({"type": "rect", "size": [1.0, 0.7], "transform": xform((0.0, 0.2, -2.5), (60.0, 0.0, 0.0)), "material": "m_roughPlastic"})
    lights = [
        {"type": "area", "color": [8.0, 7.5, 7.0], "transform": xform((-3.0, 3.5, -1.0)), "shape": {"type": "sphere", "radius": 0.5}},
        {"type": "area", "color": [6.0, 6.0, 7.0], "transform": xform((3.0, 4.0, 0.0), (90.0, 0.0, 0.0)), "shape": {"type": "plane", "size": [1.0, 0.5]}},
        {"type": "area", "color": [3.0, 4.0, 3.0], "transform": xform((6.5, 1.0, -2.0), (0.0, 30.0, 0.0)), "shape": {"type": "box", "size": [0.2, 0.4, 0.2]}},
        {"type": "point", "color": [12.0, 10.0, 8.0], "transform": xform((0.0, 3.0, -3.0))},
        {"type": "spot", "color": [40.0, 40.0, 60.0], "angle": 25.0, "transform": xform((-6.0, 3.0, -3.0), (45.0, 40.0, 0.0))},
        {"type": "background", "color": [0.15, 0.2, 0.3]},
    ]
    return {"materials": mats, "objects": objs, "lights": lights,
            "camera": {"transform": xform((0.0, 2.0, -9.0), (12.0, 0.0, 0.0)), "fieldOfView": 50.0}}


# ---- OBJ generation ---------------------------------------------------------------------------------------

class ObjWriter:
    def __init__(self):
        self.verts = []
        self.groups = {}   # material -> list of faces

    def v(self, x, y, z):
        self.verts.append((x, y, z))
        return len(self.verts)   # 1-based

    def tri(self, m, a, b, c):
        self.groups.setdefault(m, []).append((a, b, c))

    def quad(self, m, a, b, c, d):
        self.tri(m, a, b, c)
        self.tri(m, a, c, d)

    def num_tris(self):
        return sum(len(f) for f in self.groups.values())

    def write(self, path, mtl_name, materials):
        with open(path, "w") as f:
            f.write("mtllib %s\n" % mtl_name)
            f.write("".join("v %.6f %.6f %.6f\n" % v for v in self.verts))
            for m, faces in self.groups.items():
                f.write("usemtl %s\n" % m)
                f.write("".join("f %d %d %d\n" % t for t in faces))
        with open(os.path.join(os.path.dirname(path), mtl_name), "w") as f:
            for name, kd in materials.items():
                f.write("newmtl %s\nKd %.4f %.4f %.4f\nKe 0 0 0\n\n" % ((name,) + tuple(kd)))


def height_field(w, mat_name, x0, x1, z0, z1, nx, nz, fn, flip=False):
    idx = [[w.v(x0 + (x1 - x0) * i / nx, fn(x0 + (x1 - x0) * i / nx, z0 + (z1 - z0) * j / nz), z0 + (z1 - z0) * j / nz)
            for i in range(nx + 1)] for j in range(nz + 1)]
    for j in range(nz):
        for i in range(nx):
            a, b, c, d = idx[j][i], idx[j][i + 1], idx[j + 1][i + 1], idx[j + 1][i]
            if flip:
                w.quad(mat_name, a, b, c, d)
            else:
                w.quad(mat_name, a, d, c, b)


def column(w, mat_name, cx, cz, r, y0, y1, sides, rings, rng):
    rows = []
    for k in range(rings + 1):
        y = y0 + (y1 - y0) * k / rings
        rr = r * (1.0 + 0.08 * math.sin(7.0 * k / rings * math.pi)) * (1.15 if k in (0, rings) else 1.0)
        rows.append([w.v(cx + rr * math.cos(2 * math.pi * s / sides), y, cz + rr * math.sin(2 * math.pi * s / sides)) for s in range(sides)])
    for k in range(rings):
        for s in range(sides):
            a, b = rows[k][s], rows[k][(s + 1) % sides]
            c, d = rows[k + 1][(s + 1) % sides], rows[k + 1][s]
            w.quad(mat_name, a, d, c, b)


def arch(w, mat_name, x, z0, z1, y0, height, thick, segs, slices):
    # half-ring arch between two columns, in the plane x = const, extruded by 'thick'
    for t in range(slices):
        xa = x - thick / 2 + thick * t / slices
        xb = x - thick / 2 + thick * (t + 1) / slices
        prev = None
        for s in range(segs + 1):
            ang = math.pi * s / segs
            zc = (z0 + z1) / 2 - (z1 - z0) / 2 * math.cos(ang)
            yc = y0 + height * math.sin(ang)
            cur = (w.v(xa, yc, zc), w.v(xb, yc, zc), w.v(xa, yc + 0.35, zc), w.v(xb, yc + 0.35, zc))
            if prev:
                w.quad(mat_name, prev[0], prev[1], cur[1], cur[0])
                w.quad(mat_name, prev[2], cur[2], cur[3], prev[3])
            prev = cur


def drape(w, mat_name, x, z0, z1, y_top, y_bot, nz, ny, phase):
    idx = [[w.v(x + 0.25 * math.sin(phase + 9.0 * i / nz) * (j / ny), y_top + (y_bot - y_top) * j / ny, z0 + (z1 - z0) * i / nz)
            for i in range(nz + 1)] for j in range(ny + 1)]
    for j in range(ny):
        for i in range(nz):
            w.quad(mat_name, idx[j][i], idx[j][i + 1], idx[j + 1][i + 1], idx[j + 1][i])


def build_atrium(detail):
    """detail=1.0 -> ~262k triangles; smaller values scale every tessellation level down."""
    rng = random.Random(1)
    w = ObjWriter()
    L, W, H = 36.0, 14.0, 12.0   # long axis x, width z, height y
    n_long = max(4, int(200 * detail))
    n_wide = max(4, int(96 * detail))
    bumps = [(rng.uniform(-L / 2, L / 2), rng.uniform(-W / 2, W / 2), rng.uniform(0.02, 0.08)) for _ in range(40)]
    floor_fn = lambda x, z: sum(a * math.exp(-((x - bx) ** 2 + (z - bz) ** 2) * 2.0) for bx, bz, a in bumps) + 0.01 * math.sin(3 * x) * math.cos(2 * z)
    ceil_fn = lambda x, z: H + 0.6 * math.cos(z / W * math.pi) - 0.05 * math.sin(5 * x)
    height_field(w, "stone", -L / 2, L / 2, -W / 2, W / 2, n_long, n_wide, floor_fn)
    # the nave between the column rows is open to the sky (background + directional light enter there)
    height_field(w, "plaster", -L / 2, L / 2, -W / 2, -W / 4, n_long, n_wide // 2, ceil_fn, flip=True)
    height_field(w, "plaster", -L / 2, L / 2, W / 4, W / 2, n_long, n_wide // 2, ceil_fn, flip=True)
    # side walls
    n_wall_y = max(2, int(64 * detail))
    for zs, flip in ((-W / 2, False), (W / 2, True)):
        idx = [[w.v(-L / 2 + L * i / n_long, H * j / n_wall_y, zs + 0.03 * math.sin(4.0 * i / n_long * math.pi) * math.sin(3.0 * j / n_wall_y * math.pi))
                for i in range(n_long + 1)] for j in range(n_wall_y + 1)]
        for j in range(n_wall_y):
            for i in range(n_long):
                a, b, c, d = idx[j][i], idx[j][i + 1], idx[j + 1][i + 1], idx[j + 1][i]
                if flip:
                    w.quad("plaster", a, d, c, b)
                else:
                    w.quad("plaster", a, b, c, d)
    # end walls
    for xs, flip in ((-L / 2, False), (L / 2, True)):
        idx = [[w.v(xs, H * j / n_wall_y, -W / 2 + W * i / n_wide) for i in range(n_wide + 1)] for j in range(n_wall_y + 1)]
        for j in range(n_wall_y):
            for i in range(n_wide):
                a, b, c, d = idx[j][i], idx[j][i + 1], idx[j + 1][i + 1], idx[j + 1][i]
                if flip:
                    w.quad("stone", a, b, c, d)
                else:
                    w.quad("stone", a, d, c, b)
    # two rows of faceted columns + arches between them + drapes
    n_cols = 8
    sides = max(6, int(48 * detail))
    rings = max(4, int(96 * detail))
    for row_z in (-W / 4, W / 4):
        xs = [-L / 2 + L * (k + 0.5) / n_cols for k in range(n_cols)]
        for cx in xs:
            column(w, "marble", cx, row_z, 0.55, 0.0, 7.0, sides, rings, rng)
        for k in range(n_cols - 1):
            arch_slices = max(1, int(4 * detail))
            arch_segs = max(6, int(64 * detail))
            # arches run along x between neighbouring columns: reuse arch() with the x/z roles swapped via a drape-like strip
            x0, x1 = xs[k], xs[k + 1]
            for t in range(arch_slices):
                za = row_z - 0.3 + 0.6 * t / arch_slices
                zb = row_z - 0.3 + 0.6 * (t + 1) / arch_slices
                prev = None
                for s in range(arch_segs + 1):
                    ang = math.pi * s / arch_segs
                    xc = (x0 + x1) / 2 - (x1 - x0) / 2 * math.cos(ang)
                    yc = 7.0 + 1.6 * math.sin(ang)
                    cur = (w.v(xc, yc, za), w.v(xc, yc, zb), w.v(xc, yc + 0.4, za), w.v(xc, yc + 0.4, zb))
                    if prev:
                        w.quad("marble", prev[0], prev[1], cur[1], cur[0])
                        w.quad("marble", prev[2], cur[2], cur[3], prev[3])
                    prev = cur
    n_drape = max(4, int(56 * detail))
    for k in range(6):
        x = -L / 2 + L * (k + 0.5) / 6
        drape(w, "cloth_red" if k % 2 == 0 else "cloth_blue", x, -W / 4 + 0.8, W / 4 - 0.8, 9.5, 4.0 + 0.5 * (k % 3), n_drape, n_drape, 1.3 * k)
    materials = {"stone": (0.55, 0.5, 0.45), "plaster": (0.75, 0.72, 0.65), "marble": (0.8, 0.8, 0.78),
                 "cloth_red": (0.7, 0.12, 0.1), "cloth_blue": (0.1, 0.2, 0.65)}
    return w, materials


def scene_atrium(obj_rel_path):
    return {
        "materials": [
            mat("fallback", "diffuse", (0.7, 0.7, 0.7)),
            mat("chrome", "metal", (0.95, 0.93, 0.88)),
            mat("glass", "dielectric", (1.0, 1.0, 1.0), IoR=1.5),
            mat("brushed", "roughMetal", (0.9, 0.8, 0.6), roughness=0.3),
        ],
        "objects": [
            {"type": "mesh", "path": obj_rel_path, "material": "fallback"},
            {"type": "sphere", "radius": 1.1, "transform": xform((-6.0, 1.1, 0.0)), "material": "chrome"},
            {"type": "sphere", "radius": 0.9, "transform": xform((-1.0, 0.9, 1.2)), "material": "glass"},
            {"type": "sphere", "radius": 0.8, "transform": xform((4.0, 0.8, -1.0)), "material": "brushed"},
            {"type": "box", "size": [0.8, 0.5, 0.8], "transform": xform((9.0, 0.5, 0.5), (0.0, 30.0, 0.0)), "material": "chrome"},
        ],
        "lights": [
            {"type": "background", "color": [1.0, 1.5, 2.0]},
            {"type": "directional", "color": [20.0, 19.0, 18.0], "angle": 1.0, "transform": xform(None, (75.0, 20.0, 0.0))},
        ],
        "camera": {"transform": xform((-16.5, 2.2, 0.3), (2.0, 90.0, 0.0)), "fieldOfView": 60.0},
    }


# ---- BMP files + the texture scene ----------------------------------------------------------------------------

def write_bmp(path, w, h, bits, pixel, palette=None):
    """Uncompressed BMP as the reference's loader reads it (Core/Utils/BitmapBMP.cpp:46-139): 24-bit BGR, or 8-bit with
    a palette of len(palette) BGRA entries (biClrUsed > 0), or 8-bit grey without palette (biClrUsed == 0)."""
    import struct
    bpp = 3 if bits == 24 else 1
    stride = (w * bpp + 3) & ~3
    rows = bytearray()
    for y in range(h):
        row = bytearray()
        for x in range(w):
            v = pixel(x, y)
            row += bytes(v) if bits == 24 else bytes([v])
        row += bytes(stride - len(row))
        rows += row
    pal = b"".join(bytes(c) for c in (palette or []))
    off = 14 + 40 + len(pal)
    with open(path, "wb") as f:
        f.write(struct.pack("<HIHHI", 0x4D42, off + len(rows), 0, 0, off))
        f.write(struct.pack("<IIIHHIIIIII", 40, w, h, 1, bits, 0, len(rows), 2835, 2835, len(palette or []), 0))
        f.write(pal)
        f.write(rows)


def write_texture_files(out_dir):
    tex = os.path.join(out_dir, "TEXTURES")
    os.makedirs(tex, exist_ok=True)
    # (b, g, r) per pixel
    write_bmp(os.path.join(tex, "tiles.bmp"), 16, 8, 24, lambda x, y: ((x ^ y) * 37 % 256, y * 32 % 256, x * 16 % 256 if (x + y) % 3 else 255))
    write_bmp(os.path.join(tex, "normal.bmp"), 8, 8, 24,
              lambda x, y: (255, int(128 + 70 * math.cos(y * 0.9 + x * 0.3)), int(128 + 70 * math.sin(x * 0.8))))
    write_bmp(os.path.join(tex, "env.bmp"), 32, 16, 24,
              lambda x, y: (min(255, 90 + 10 * y), min(255, 60 + 8 * y + (150 if (x, y) == (20, 5) else 0)), min(255, 40 + 4 * x + (200 if (x, y) == (20, 5) else 0))))
    palette = [((i * 53) % 256, (i * 97 + 30) % 256, (255 - i * 16) % 256, (i * 17) % 256) for i in range(16)]
    write_bmp(os.path.join(tex, "pal.bmp"), 13, 5, 8, lambda x, y: (x * 3 + y * 5) % 16, palette)
    write_bmp(os.path.join(tex, "grey.bmp"), 6, 6, 8, lambda x, y: (x * 40 + y * 9) % 256)
    # bokeh bitmap: a ring with a few black texels (zero importance) and a bright spot
    write_bmp(os.path.join(tex, "bokeh.bmp"), 11, 9, 24,
              lambda x, y: (0, 0, 0) if (x - 5) ** 2 + (y - 4) ** 2 > 20 or (x - 5) ** 2 + (y - 4) ** 2 < 3 else ((255, 255, 255) if (x, y) == (7, 5) else (60 + 10 * x, 90, 40 + 15 * y)))


def write_textured_mesh(out_dir):
    """A folded strip with texture coordinates; the .mtl binds a diffuse texture and a normal map by file name."""
    n = 6
    lines = ["mtllib textured.mtl"]
    for i in range(n + 1):
        x = -1.5 + 3.0 * i / n
        for j, z in enumerate((-0.6, 0.6)):
            lines.append("v %.6f %.6f %.6f" % (x, 0.25 * math.sin(i * 1.3) + 0.1 * j, z))
    for i in range(n + 1):
        for j in range(2):
            lines.append("vt %.6f %.6f" % (1.7 * i / n - 0.2, 0.9 * j + 0.05 * i))
    lines.append("usemtl cloth")
    for i in range(n):
        a, b, c, d = 2 * i + 1, 2 * i + 2, 2 * i + 4, 2 * i + 3
        lines.append("f %d/%d %d/%d %d/%d" % (a, a, b, b, c, c))
        lines.append("f %d/%d %d/%d %d/%d" % (a, a, c, c, d, d))
    with open(os.path.join(out_dir, "MODELS", "textured.obj"), "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.join(out_dir, "MODELS", "textured.mtl"), "w") as f:
        f.write("newmtl cloth\nKd 0.9 0.8 0.7\nKe 0 0 0\nmap_Kd ../TEXTURES/tiles.bmp\nnorm ../TEXTURES/normal.bmp\n\n")


def scene_textures():
    return {
        "textures": [
            {"name": "checker", "type": "checkerboard", "colorA": [0.9, 0.9, 0.9], "colorB": [0.1, 0.1, 0.3]},
            {"name": "clouds", "type": "noise", "colorA": [0.0, 0.0, 0.0], "colorB": [1.0, 1.0, 1.0], "octaves": 4},
            {"name": "tiles", "type": "bitmap", "path": "TEXTURES/tiles.bmp"},
            {"name": "blend", "type": "mix", "textureA": "checker", "textureB": "tiles", "weight": "clouds"},
            {"name": "blend2", "type": "mix", "textureA": "blend", "textureB": "TEXTURES/pal.bmp", "weight": "checker"},
        ],
        "materials": [
            mat("floor", "diffuse", (1.0, 1.0, 1.0), baseColorTexture="blend"),
            mat("ball", "roughMetal", (0.95, 0.9, 0.8), roughness=0.8, roughnessTexture="clouds", baseColorTexture="TEXTURES/pal.bmp"),
            mat("crate", "diffuse", (0.9, 0.9, 0.9), baseColorTexture="tiles", normalMap="TEXTURES/normal.bmp", normalMapStrength=0.7),
            mat("glow", "diffuse", (0.2, 0.2, 0.2), emissionColor=[2.0, 1.5, 1.0], emissionTexture="checker"),
            mat("coat", "roughPlastic", (0.8, 0.8, 0.8), roughness=0.4, metalness=0.9, metalnessTexture="TEXTURES/grey.bmp", baseColorTexture="blend2"),
            mat("ghost", "diffuse", (0.3, 0.7, 0.4), baseColorTexture="TEXTURES/missing.bmp"),
            mat("meshfallback", "diffuse", (0.7, 0.7, 0.7)),
        ],
        "objects": [
            {"type": "rect", "size": [6.0, 6.0], "textureScale": [0.25, 0.4], "transform": xform((0.0, -1.0, 0.0), (-90.0, 0.0, 0.0)), "material": "floor"},
            {"type": "sphere", "radius": 0.8, "transform": xform((-2.2, -0.2, 0.3)), "material": "ball"},
            {"type": "box", "size": [0.6, 0.6, 0.6], "transform": xform((0.0, -0.4, 0.8), (0.0, 30.0, 0.0)), "material": "crate"},
            {"type": "sphere", "radius": 0.35, "transform": xform((1.2, -0.65, -0.9)), "material": "glow"},
            {"type": "box", "size": [0.5, 0.8, 0.4], "transform": xform((2.3, -0.2, 0.6), (0.0, -20.0, 0.0)), "material": "coat"},
            {"type": "sphere", "radius": 0.4, "transform": xform((-0.9, -0.6, -1.2)), "material": "ghost"},
            {"type": "mesh", "path": "MODELS/textured.obj", "transform": xform((0.3, 0.6, 1.6), (20.0, 0.0, 0.0)), "material": "meshfallback"},
        ],
        "lights": [
            {"type": "background", "color": [1.0, 1.0, 1.0], "texture": "TEXTURES/env.bmp"},
            {"type": "area", "color": [9.0, 9.0, 8.0], "transform": xform((0.0, 3.0, -0.5), (90.0, 0.0, 0.0)), "shape": {"type": "plane", "size": [0.8, 0.8]}},
        ],
        "camera": {"transform": xform((0.0, 1.2, -5.5), (14.0, 0.0, 0.0)), "fieldOfView": 55.0},
    }


def scene_caustics():
    """BASELINE config 5 in miniature: a rough-glass sphere and a glass block between a small area light and a diffuse
    floor (caustics: the paths the bidirectional renderer exists for), plus a metal sphere."""
    return {
        "materials": [
            mat("floor", "diffuse", (0.8, 0.8, 0.8)),
            mat("wall", "diffuse", (0.7, 0.3, 0.3)),
            mat("roughGlass", "roughDielectric", (1.0, 1.0, 1.0), roughness=0.08, IoR=1.5),
            mat("glass", "dielectric", (0.95, 1.0, 0.95), IoR=1.45),
            mat("steel", "roughMetal", (0.9, 0.9, 0.9), roughness=0.2),
        ],
        "objects": [
            {"type": "box", "size": [4.0, 0.1, 4.0], "transform": xform((0.0, -1.1, 0.0)), "material": "floor"},
            {"type": "box", "size": [4.0, 2.0, 0.1], "transform": xform((0.0, 0.9, 3.0)), "material": "wall"},
            {"type": "sphere", "radius": 0.6, "transform": xform((-0.9, -0.4, 0.2)), "material": "roughGlass"},
            {"type": "box", "size": [0.35, 0.5, 0.35], "transform": xform((0.9, -0.5, 0.0), (0.0, 35.0, 0.0)), "material": "glass"},
            {"type": "sphere", "radius": 0.35, "transform": xform((0.1, -0.65, -1.0)), "material": "steel"},
        ],
        "lights": [{"type": "area", "color": [60.0, 55.0, 50.0], "transform": xform((0.0, 2.6, 0.0), (90.0, 0.0, 0.0)),
                    "shape": {"type": "plane", "size": [0.25, 0.25]}}],
        "camera": {"transform": xform((0.0, 0.9, -5.0), (14.0, 0.0, 0.0)), "fieldOfView": 48.0},
    }


def write_all(out_dir, sponza=False):
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(os.path.join(out_dir, "MODELS"), exist_ok=True)

    def dump(name, scene):
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(scene, f, indent=1)

    dump("c1_sphere_area", scene_c1())
    dump("cornell", scene_cornell())
    dump("materials", scene_materials())
    dump("caustics", scene_caustics())
    dof = scene_cornell()
    dof["camera"].update({"enableDOF": True, "aperture": 0.15, "focalPlaneDistance": 3.5})    # circle bokeh, Camera.cpp:100-116
    dump("cornell_dof", dof)
    bokeh = scene_cornell()
    bokeh["camera"].update({"enableDOF": True, "aperture": 0.25, "focalPlaneDistance": 3.5, "bokehTexture": "TEXTURES/bokeh.bmp"})   # Camera.cpp:208-213
    dump("cornell_bokeh", bokeh)
    write_texture_files(out_dir)
    write_textured_mesh(out_dir)
    dump("textures", scene_textures())

    w, mats = build_atrium(0.09)
    w.write(os.path.join(out_dir, "MODELS", "mesh_small.obj"), "mesh_small.mtl", mats)
    dump("mesh_small", scene_atrium("MODELS/mesh_small.obj"))
    info = {"mesh_small_tris": w.num_tris()}
    if sponza:
        w, mats = build_atrium(0.864)
        w.write(os.path.join(out_dir, "MODELS", "sponza_class.obj"), "sponza_class.mtl", mats)
        dump("sponza_class", scene_atrium("MODELS/sponza_class.obj"))
        info["sponza_class_tris"] = w.num_tris()
    return info


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.abspath(__file__))
    print(json.dumps(write_all(out, "--sponza" in sys.argv)))
