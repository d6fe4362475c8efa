#!/bin/bash
# A/B timing on ONE GPU box: one short bench per argument, prints fused / field-stage ms.  An argument of the form
# lib:<path> switches the library build (NGM_B200_LIB) for the following runs; any other argument is exported as
# NGM_TC_FLAGS (a free experiment knob a development build may read) and labels the run.
TAG=${1:-sweep}; shift
mkdir -p gpurun_out
# an argument of the form lib:<path> switches the library for the following values
for fl in "$@"; do
  case "$fl" in lib:*) export NGM_B200_LIB="$PWD/${fl#lib:}"; echo "library: $NGM_B200_LIB"; continue;; esac
  NGM_TC_FLAGS=$fl timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/${TAG}_flags$fl.json 2> gpurun_out/${TAG}_flags$fl.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_flags$fl.json"))
    st = d["roofline_stages"]
    print("flags=$fl  step %.3f ms  fused %.3f ms  field %.3f ms  e2e %.3f ms" % (d["ms_per_step"], st["render_fused"]["ms"], st["field_mlp"]["ms"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("flags=$fl FAILED", e); print(open("gpurun_out/${TAG}_flags$fl.err").read()[-1500:])
PY
done
