#!/bin/bash
# A/B runs of bench.py (weak-scaling leg only) for library variants / tuning knobs on one GPU box.
#   tools/ab_bench.sh <tag> [ENV=VAL ...] -- [extra bench args]
# Lines go to gpurun_out/ab_<tag>.json; the summary (value, e2e, kernel split) is printed.
tag=$1; shift
envs=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do envs+=("$1"); shift; done
[ "$1" == "--" ] && shift
env "${envs[@]}" timeout 120 python bench.py --no-configs --frame-spp 0 --no-cpu-baseline --steps 6 "$@" > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{tag}.json"))
    k = d["kernel_split_profiled"]
    print(f"{tag:28s} value {d['value']:8.1f}  e2e {d['e2e']['value']:8.1f}  world {k.get('extend_world_ms', k.get('extend_ms', 0)):7.1f} inst {k.get('extend_instances_ms', 0):6.1f} "
          f"finish {k['finish_shade_ms']:7.1f} deferred {k['deferred_shade_ms']:5.1f}  pairs/ray {d['roofline']['per_ray']['aabb_tests'] / 2:5.2f}")
except Exception as e:
    print(tag, "FAILED", e, open(f"gpurun_out/ab_{tag}.err").read()[-400:])
PY
