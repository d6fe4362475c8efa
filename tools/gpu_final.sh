#!/bin/bash
# One short gpurun call: the newest GPU tests, then (unless "quick") the rest of the GPU suite, and the ncu launch
# list of the training-side kernels.  Outputs under gpurun_out/.
TAG=${1:-r1f}
NEW="tests/test_gpu_targets.py::test_get_observed_fields_golden_and_random tests/test_gpu_fullsize.py::test_c5_loop_closure_rerender_invariance"
mkdir -p gpurun_out
timeout 300 python -m pytest $NEW -x -q > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "new tests rc=$?"; tail -15 gpurun_out/${TAG}_pytest_new.log
if [ "$2" != "quick" ]; then
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
fi
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'adam|target|observed' --csv --log-file gpurun_out/${TAG}_train_kernels.csv python tools/prof_train.py > gpurun_out/${TAG}_prof_train.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_prof_train.log
grep -c "adam\|target\|observed" gpurun_out/${TAG}_train_kernels.csv
