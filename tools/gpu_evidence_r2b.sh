#!/bin/bash
# Round-2 final evidence run (one GPU): tests, the bench line, the ncu launch list of the bench command and full
# captures of the kernels that changed late in the round.  Outputs under gpurun_out/ (tools/ncu_summary.py -> profiles/).
set -x
B="python bench.py --no-extras --no-cpu-baseline"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/r2_bench_fp16_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
NCU="ncu --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2b_launches_fp16.csv $B --steps 2 --warmup 3 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:tc_kernel -s 4 -c 1 -o gpurun_out/r2b_fused -f $B --steps 2 --warmup 3 > /dev/null 2>&1
timeout 300 $NCU --set full -k regex:"permuto_rows" -c 1 -o gpurun_out/r2b_permuto -f python tools/knn_timeline.py "permuto_1x32 (reference default field)" fp16 vmap > /dev/null 2>&1
timeout 300 $NCU --set full -k regex:"knn_assign|knn_scatter|knn_blend" -c 3 -o gpurun_out/r2b_knn -f python tools/bench_knn.py --only nerf8_4x128:fp16 > /dev/null 2>&1
timeout 300 python tools/bench_knn.py > gpurun_out/r2b_knn.jsonl 2>&1
ls -la gpurun_out/r2b_*
