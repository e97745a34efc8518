#!/bin/bash
# Builds every flavour of the library in-tree: shipped, DEBUG_CHECKS (device-side self checks) and EXPERIMENTS (A/B knobs).
set -e
cd "$(dirname "$0")/../raytrace2_b200/csrc"
make -j"$(nproc)" 2>&1 | grep -E "error|Error" || true
make -j"$(nproc)" DEBUG_CHECKS=1 2>&1 | grep -E "error|Error" || true
make -j"$(nproc)" EXPERIMENTS=1 2>&1 | grep -E "error|Error" || true
ls -la ../lib
