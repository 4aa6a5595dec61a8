#!/bin/bash
# per-phase timeline of the persistent kernel (TIMELINE build), then restore the production build
set -x
B2SVM_TIMELINE=1 python -m pysvm_b200.build --force > /dev/null 2>&1
B2SVM_TIMELINE=1 python scripts/gpu_timeline.py 200000:128:f32 25000:128:f32 4096:128:f32 > gpurun_out/r2_timeline.log 2>&1
CACHE_BYTES=100000000000 FIRST_ITERS=3000 B2SVM_TIMELINE=1 python scripts/gpu_timeline.py 20000:128:f32 > gpurun_out/r2_timeline_warm.log 2>&1
python -m pysvm_b200.build --force > /dev/null 2>&1
