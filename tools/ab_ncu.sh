#!/bin/bash
# per-build kernel times under ncu (serialised, one launch each) for the A/B libraries of tools/ab_bench.sh
for v in A B; do
  SDB_LIB_PATH=$PWD/slenderobjdet_b200/csrc/build/ab/lib$v.so timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum --clock-control none -k regex:"dx_gather|dcn_bwd_data" -s 2 -c 2 --csv python tools/prof_step.py --steps 2 2>/dev/null | grep -E "dx_gather|dcn_bwd_data" | awk -F'","' -v t=$v '{print t, substr($5,1,40), $(NF-2), $(NF)}'
done
