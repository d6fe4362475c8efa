#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the in-kernel phase trace, the ncu launch list
# of the bench command and one `ncu --set full` capture of every kernel.  Outputs under gpurun_out/.
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json
for fl in $NGM_FLAG_SWEEP; do
  NGM_TC_FLAGS=$fl timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_flags$fl.json 2>> gpurun_out/${TAG}_bench.err
  echo "flags=$fl: $(python -c "import json,sys; d=json.load(open('gpurun_out/${TAG}_bench_flags$fl.json')); print(d['ms_per_step'], d['roofline_stages']['field_mlp']['ms'])")"
done
timeout 200 python tools/tc_trace.py > gpurun_out/${TAG}_trace.txt 2>&1; echo "trace rc=$?"
if [ "$2" != "noprof" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_fp16.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${TAG}_prof \
    python tools/prof_stages.py > gpurun_out/${TAG}_prof.log 2>&1; echo "ncu full rc=$?"
fi
