#!/bin/bash
# short bench line + kernel table: profiles/benchshort.sh <tag> [extra bench args]
tag=$1; shift
python bench.py --steps 20 --warmup 20 --no-cpu-baseline --no-e2e --no-extra "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("$tag", round(d["ms_per_step"],4), "ms/gen", round(d["value"]/1e9,2), "G/s")
for k,v in d["kernels"].items():
    if v["share"]>0.01: print("   ", k, v["ms_per_launch"], v["share"])
PY
tail -2 gpurun_out/bench_$tag.err
