#!/bin/bash
# Round-2 evidence for the stage pipeline (one GPU): launch list of the bench command and a full ncu capture of the
# dominant kernel (tcgen05 field stage) and of the compositor on the headline workload.
set -x
B="python bench.py --no-extras --no-cpu-baseline"
NCU="ncu --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2c_launches_fp16.csv $B --steps 2 --warmup 3 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"tc_kernel|composite_staged|sample_rays" -s 9 -c 3 -o gpurun_out/r2c_stages -f $B --steps 2 --warmup 3 > /dev/null 2>&1
timeout 300 python bench.py > gpurun_out/r2_bench_fp16_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
timeout 300 python tools/fused_vs_staged.py > gpurun_out/r2_fused_vs_staged.jsonl 2>/dev/null
ls -la gpurun_out/r2c_*
