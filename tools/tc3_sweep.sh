#!/bin/bash
# field-stage kernel timing: two-slot tc_kernel vs the three-slot kernel with 1/2/3 tiles in flight
run() { python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'field_mlp ms', round(d['roofline_stages']['field_mlp']['ms'],3))"; }
run "two-slot"
for n in 1 2 3; do NGM_TC3=1 NGM_TC3_SLOTS=$n run "tc3 slots=$n"; done
