"""pytest configuration: the `gpu` marker and shared fixtures.

CPU tests (`-m "not gpu"`) cover the oracle against the golden vectors / the real reference, the host scene compiler, the
image writer, the multi-process host logic (gloo) and that the C-ABI library loads and exports every declared symbol.
GPU tests (`-m gpu`) are the parity tests proper and call the CUDA path through the C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")

# BASELINE.json configs 1-4 (+ nested-transform scenes that exercise instance chains)
CURRENT_SCENES = ["cornell_original_test", "cornell_box_scene_graph", "cornell_box4", "cornell_volume_10000_samples",
                  "book2_final_scene_10000_samples"]
LEGACY_SCENES = ["final_render_book_1", "final_render_scene_blur", "scene2", "cornell_box2", "light_scene1", "checker_test"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def scene_path(name: str) -> str:
    return os.path.join(DATA, name + ".json")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if needed) and load libraytrace2_b200.so — fails loudly when it cannot be built."""
    import raytrace2_b200 as rt
    if not os.path.exists(rt.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return rt.load_library()


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import rt_oracle
    rt_oracle.lib()
    return rt_oracle


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle import ref_oracle
    if not ref_oracle.available():
        pytest.skip("oracle/_ref/libref_oracle.so not built (needs /root/reference)")
    return ref_oracle


def isotropic_lookup(name: str):
    """Boolean table over material indices (+ a trailing False for -1 = miss): True for isotropic materials, i.e. hits
    produced by a constant medium, whose Hit() is stochastic (ConstantMedium.cpp:42)."""
    import numpy as np
    import raytrace2_b200 as rt
    mats = rt.Scene.load(scene_path(name)).materials()
    return np.array([int(m["type"]) == 5 for m in mats] + [False])
