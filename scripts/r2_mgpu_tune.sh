#!/bin/bash
# poll pacing variants at N GPUs: one short bench line per setting
N=${1:-2}
shift
run() {
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 --no-full-fit --no-secondary --no-cpu-baseline > gpurun_out/tune.json 2> gpurun_out/tune.err
  python - "$*" <<PY
import json, sys
try:
    d=json.loads([l for l in open("gpurun_out/tune.json") if l.startswith("{")][-1])
    print(sys.argv[1], "-> value", round(d["value"]), "us/it", round(d["us_per_iteration"],3), "e2e", round(d["e2e"]["value"]), "sha", d["parity"]["alpha_sha256"][:16])
except Exception as e:
    print(sys.argv[1], "no bench line", e)
PY
}
for setting in "$@"; do run $setting; done
