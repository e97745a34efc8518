#!/usr/bin/env python3
"""Re-serialise the reference's scene / camera JSON files (inputs, not code) into data/ in canonical minified form.

The BASELINE configs name concrete scene files of the reference repository (data/*.json); the GPU box has no
/root/reference, so the inputs travel with this repo.  Values are untouched (json round-trip keeps every number's
shortest repr); only whitespace changes.  Run in the build container:  python tools/import_scenes.py
"""
import glob
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")


def main():
    os.makedirs(OUT, exist_ok=True)
    for path in sorted(glob.glob(os.path.join(REF, "*.json"))):
        with open(path) as f:
            doc = json.load(f)
        out = os.path.join(OUT, os.path.basename(path))
        with open(out, "w") as f:
            json.dump(doc, f, separators=(",", ":"))
            f.write("\n")
        print(f"{os.path.basename(path)}: {os.path.getsize(path)} -> {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
