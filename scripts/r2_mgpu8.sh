#!/bin/bash
# 8-GPU measurement: bench (default plan), bench with all SMs per GPU, per-phase timeline
N=${1:-8}
run_bench() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 --no-full-fit --no-secondary > gpurun_out/r2_bench_n${N}$1.json 2> gpurun_out/r2_bench_n${N}$1.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_n${N}$1.json") if l.startswith("{")][-1])
    print("N=$N$1 value", round(d["value"]), "us/it", round(d["us_per_iteration"],3), "e2e", round(d["e2e"]["value"]), "sha", d["parity"]["alpha_sha256"][:16], "e2e sha", d["e2e"]["alpha_sha256"][:16])
except Exception as e:
    print("no bench line", e)
PY
}
run_bench ""
B2SVM_MIN_TILES_PER_CTA=5 run_bench "_allsm"
B2SVM_TIMELINE=1 python -m pysvm_b200.build --force > /dev/null 2>&1
B2SVM_TIMELINE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29535 scripts/mgpu_timeline.py > gpurun_out/r2_mgpu_tl_n$N.log 2>&1
grep -A13 "^rank 0" gpurun_out/r2_mgpu_tl_n$N.log
python -m pysvm_b200.build --force > /dev/null 2>&1
