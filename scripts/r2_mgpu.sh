#!/bin/bash
# multi-GPU bring-up on an N-GPU box: sharded parity worker, bench at N, per-phase timeline (TIMELINE build)
N=${1:-2}
set -x
if [ "$N" = "2" ]; then
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 tests/_sharded_worker.py > gpurun_out/r2_sharded_worker.log 2>&1; echo "worker rc=$?"; grep "sharded worker ok\|Error\|error\|assert" gpurun_out/r2_sharded_worker.log | head
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 --no-full-fit --no-secondary > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_n$N.json") if l.startswith("{")][-1])
    print("N=$N value", d["value"], "us/it", d["us_per_iteration"], "e2e", d["e2e"]["value"], "parity", d["parity"]["alpha_sha256"][:16], d["parity"]["next_pair"], "e2e sha", d["e2e"]["alpha_sha256"][:16])
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/r2_bench_n$N.err
B2SVM_TIMELINE=1 python -m pysvm_b200.build --force > /dev/null 2>&1
B2SVM_TIMELINE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29535 scripts/mgpu_timeline.py > gpurun_out/r2_mgpu_tl_n$N.log 2>&1
grep -A13 "^rank 0" gpurun_out/r2_mgpu_tl_n$N.log
python -m pysvm_b200.build --force > /dev/null 2>&1
