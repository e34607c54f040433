#!/bin/bash
# ncu --set full of one instance of every ORB kernel (first launch of each = pyramid level 0 / 1)
mkdir -p gpurun_out
for k in distribute_kernel cell_nms_kernel blur7_kernel fast_score_kernel resize_kernel orient_describe_kernel; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 1 -o gpurun_out/ncu_orb_$k -f python tools/prof_orb.py > gpurun_out/ncu_orb_$k.log 2>&1
done
ls -la gpurun_out | grep ncu_orb_
