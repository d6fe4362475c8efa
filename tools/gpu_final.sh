#!/bin/bash
# One short gpurun call: the newest GPU tests first (fail fast), then the rest of the GPU suite, the optimizer
# micro-benchmark and the bench line.  Outputs under gpurun_out/.
TAG=${1:-r1f}
NEW="tests/test_gpu_targets.py tests/test_checkpoint.py"
mkdir -p gpurun_out
timeout 300 python -m pytest $NEW -m gpu -x -q > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "new tests rc=$?"; tail -15 gpurun_out/${TAG}_pytest_new.log
DESEL=""; for t in $NEW; do DESEL="$DESEL --deselect $t"; done
timeout 600 python -m pytest tests -m gpu -x -q $DESEL > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
if [ "$2" != "nobench" ]; then
timeout 120 python tools/bench_adam.py > gpurun_out/${TAG}_bench_adam.json 2> gpurun_out/${TAG}_bench_adam.err; echo "adam rc=$?"; cat gpurun_out/${TAG}_bench_adam.json; tail -3 gpurun_out/${TAG}_bench_adam.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
fi
