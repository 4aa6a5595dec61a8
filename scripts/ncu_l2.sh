#!/bin/bash
# DRAM traffic and L2 hit rate of the persistent kernel for a few L2 residency plans (B2SVM_L2_PIN_MB / _SKIP)
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
for cfg in "none" "56:0" "80:0" "110:0" "88:12"; do
  if [ "$cfg" = "none" ]; then unset B2SVM_L2_PIN_MB B2SVM_L2_PIN_SKIP; else export B2SVM_L2_PIN_MB=${cfg%%:*} B2SVM_L2_PIN_SKIP=${cfg##*:}; fi
  echo "== plan $cfg"
  ITERS=400 timeout 200 ncu --metrics $M --clock-control none -k regex:smo_persistent -s 1 -c 1 --csv python scripts/ncu_target.py 2>/dev/null | grep -E "dram__|lts__|gpu__time" | awk -F'","' '{print "   " $(NF-2) " = " $NF " " $(NF-1)}'
done
