#!/bin/bash
# usage: sweep.sh tag "ENV=.. ENV=.." ; runs cfg2 128 MiB bench without verification
tag=$1; shift
env "$@" timeout 300 python bench.py --bytes 134217728 --steps 3 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sw_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "value %.4f e2e %.4f ms %.1f matches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["matches_per_step"]), d.get("states_per_s",{}).get("visited_by_kernel"))
except Exception as e:
    print("$tag FAILED", e)
PY
