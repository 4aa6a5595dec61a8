#!/bin/bash
# one clean bench line at N GPUs (value leg, e2e leg, converged fit, C2) -> gpurun_out/r2_bench_nN.json
N=${1:-2}
timeout ${2:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_n$N.json") if l.startswith("{")][-1])
    print("N=$N value", round(d["value"]), "us/it", round(d["us_per_iteration"],3), "e2e", round(d["e2e"]["value"]), "sha", d["parity"]["alpha_sha256"][:16], "ctas", d.get("config",{}).get("parallelism"))
    print("fit", d.get("fit"))
    print("sec", d.get("secondary"))
except Exception as e:
    print("no bench line", e)
PY
tail -2 gpurun_out/r2_bench_n$N.err
