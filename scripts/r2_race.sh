#!/bin/bash
B2SVM_TIMELINE=1 python -m pysvm_b200.build --force > /dev/null 2>&1
WORLD=2 python scripts/gpu_race_hunt.py 2>&1 | tail -60

python -m pysvm_b200.build --force > /dev/null 2>&1
