#!/bin/bash
# Round-2 evidence run (one GPU): A/B of the fp16-accumulator variant, ncu launch list + full captures of the kernels
# profiles/README.md quotes.  Outputs under gpurun_out/ (summarised into profiles/ with tools/ncu_summary.py).
set -x
B="python bench.py --no-extras --no-cpu-baseline"
for v in 0 1; do NGM_TC_ACC16=$v $B --steps 40 --warmup 5 2>/dev/null > gpurun_out/r2_bench_acc16_$v.json; done
python -m pytest tests/test_gpu_packed_weights.py -q 2>&1 | tail -4
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2_launches_fp16.csv $B --steps 2 --warmup 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:tc_kernel -s 4 -c 1 -o gpurun_out/r2_fused -f $B --steps 2 --warmup 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"bwd_kernel" -s 3 -c 3 -o gpurun_out/r2_train -f python tools/bench_train.py --profile c4 > /dev/null 2>&1
$NCU --set full -k regex:"permuto_rows|feature_rows" -c 1 -o gpurun_out/r2_permuto -f python tools/bench_variants.py > gpurun_out/r2_variants_under_ncu.log 2>&1
$NCU --set full -k regex:"knn_" -c 4 -o gpurun_out/r2_knn -f python tools/bench_knn.py > gpurun_out/r2_knn_under_ncu.log 2>&1
$NCU --set full -k regex:"sample_rays|composite_staged" -c 2 -o gpurun_out/r2_stages -f python tools/prof_stages.py > /dev/null 2>&1
python tools/mapping_loop.py --frames 150 > gpurun_out/r2_mapping_loop.jsonl 2>&1
python tools/bench_variants.py > gpurun_out/r2_variants.jsonl 2>&1
python tools/bench_knn.py > gpurun_out/r2_knn.jsonl 2>&1
ls -la gpurun_out/*.ncu-rep
