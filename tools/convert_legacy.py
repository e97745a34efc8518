#!/usr/bin/env python3
"""Legacy -> current scene format.

13 of the reference's 22 scene files (`data/scene2.json`, `final_render_book_1.json`, ...) are in a format its loader at
HEAD no longer reads (`"primitives": {"spheres": [...], "quads": [...]}`, materials with "id", no "scene" array, camera
sometimes absent; SURVEY Appendix B: HEAD throws nlohmann type_error 302 / 306 on them).  This tool rewrites such a file into
the CURRENT format (Serialize.cpp:199-360) under exactly the rules of our legacy adapter (csrc/host/scene_host.cpp,
oracle/rt_oracle.py): spheres, then quads, then boxes, each a top-level scene node; `material_id` resolved through the
materials' "id"; a named camera stays a name, an absent camera becomes data/cam1.json's contents for `final_render_*` and the
loader's defaults otherwise.  The converted file loads in the UNMODIFIED reference, which is how tests/golden/make_golden.py
pins the legacy scenes (BASELINE config 2 among them) to the reference's own Hit() and render.

    python tools/convert_legacy.py data/final_render_book_1.json /tmp/book1_current.json
"""
from __future__ import annotations

import json
import os
import sys


def is_legacy(doc: dict) -> bool:
    return isinstance(doc.get("primitives"), dict)


def convert(doc: dict, base_name: str, data_dir: str) -> dict:
    if not is_legacy(doc):
        return doc
    out: dict = {}
    cam = doc.get("camera")
    if isinstance(cam, (dict, str)):
        out["camera"] = cam
    elif base_name.startswith("final_render"):
        with open(os.path.join(data_dir, "cam1.json")) as f:
            out["camera"] = json.load(f)
    else:
        out["camera"] = {}
    if "background_color" in doc:
        out["background_color"] = doc["background_color"]
    if isinstance(doc.get("textures"), list):
        out["textures"] = doc["textures"]
    mats, ids = [], {}
    for i, m in enumerate(doc["materials"]):
        ids[m.get("id", i)] = i
        m2 = {k: v for k, v in m.items() if k != "id"}
        if not m2.get("type"):
            if "tex_idx" in m2:
                m2["type"] = "texture"
            else:
                raise ValueError("material type field empty")
        mats.append(m2)
    out["materials"] = mats

    def mat_of(p):
        mid = p.get("material_id", p.get("material", 0))
        return ids.get(mid, mid)

    prims = []
    pr = doc["primitives"]
    for p in pr.get("spheres", []):
        q = {"type": "sphere", "center": p.get("center", [0, 0, 0]), "radius": p.get("radius", 0.5), "material": mat_of(p)}
        if "displacement" in p:
            q["displacement"] = p["displacement"]
        if "constant_medium" in p:
            q["constant_medium"] = p["constant_medium"]
        prims.append(q)
    for p in pr.get("quads", []):
        q = {"type": "quad", "q": p.get("q", [0, 0, 0]), "u": p.get("u", [1, 0, 0]), "v": p.get("v", [0, 0, 1]), "material": mat_of(p)}
        if "constant_medium" in p:
            q["constant_medium"] = p["constant_medium"]
        prims.append(q)
    for p in pr.get("boxes", []):
        q = {"type": "box", "a": p.get("a", [0, 0, 0]), "b": p.get("b", [1, 1, 1]), "material": mat_of(p)}
        if "constant_medium" in p:
            q["constant_medium"] = p["constant_medium"]
        prims.append(q)
    out["primitives"] = prims
    out["scene"] = [{"primitive": i} for i in range(len(prims))]
    return out


def convert_file(src: str, dst: str, data_dir: str | None = None) -> None:
    with open(src) as f:
        doc = json.load(f)
    base = os.path.basename(src)
    res = convert(doc, base, data_dir or os.path.dirname(os.path.abspath(src)))
    with open(dst, "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    convert_file(sys.argv[1], sys.argv[2])
