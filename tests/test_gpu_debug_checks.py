"""GPU: the library's own device-side checks (compute-sanitizer is closed on this GPU pool — gpurun answers "compute-sanitizer is
closed on this pool and stays closed", profiles/r02_sanitizer.md).  libraytrace2_b200_dbg.so (`make DEBUG_CHECKS=1`, built by
__graft_entry__.build()) verifies every data-dependent index inside the kernels and poisons the wavefront buffers with NaNs;
tools/sanitize_target.py drives every kernel family through it (all three instance walks, flat / wide / LBVH trees, fused and
per-bin shading, media pass, deferred noise shading, resolve kernels, multi-GPU handle when 2 GPUs are visible).  All violation
counters must be zero and no pixel may be NaN."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

DBG = os.path.join(ROOT, "raytrace2_b200", "lib", "libraytrace2_b200_dbg.so")


def _run(lib_path):
    env = dict(os.environ)
    if lib_path:
        env["RT2_LIB_PATH"] = lib_path
    else:
        env.pop("RT2_LIB_PATH", None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_target.py")], capture_output=True, text=True, timeout=900,
                       env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("DEBUG_CHECKS ")][-1]
    return json.loads(line[len("DEBUG_CHECKS "):]), p.stdout


def test_device_side_checks_find_no_violation(native_lib):
    if not os.path.exists(DBG):
        import __graft_entry__
        __graft_entry__.build_debug_checks()
    chk, out = _run(DBG)
    assert chk["enabled"] is True, "the _dbg library must be built with -DRT2_DEBUG_CHECKS"
    bad = {k: v for k, v in chk["violations"].items() if v}
    assert not bad, f"device-side index checks failed: {bad}"
    assert chk["nan_pixels"] == 0, "a kernel read wavefront state that nobody wrote (poisoned buffers)"
    assert "book2/unified" in out and "book2/split" in out and "book2/inline" in out and "synthetic/" in out


def test_shipped_library_has_the_checks_compiled_out(native_lib):
    chk, _ = _run(None)
    assert chk["enabled"] is False and not any(chk["violations"].values()) and chk["nan_pixels"] == 0
