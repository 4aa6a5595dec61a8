#!/bin/bash
# timeline of the cached-column (warm) iterations at full size: 190k iterations fill the cache first
B2SVM_TIMELINE=1 python -m pysvm_b200.build --force > /dev/null 2>&1
PER_EPOCH=1 CACHE_BYTES=120000000000 FIRST_ITERS=190000 B2SVM_TIMELINE=1 python scripts/gpu_timeline.py 200000:128:f32 > gpurun_out/r2_timeline_warm_full.log 2>&1
python -m pysvm_b200.build --force > /dev/null 2>&1
tail -90 gpurun_out/r2_timeline_warm_full.log
