#!/bin/bash
# per-variant time and L1/L2 hit rates of the grad_input gather (ncu, one launch each)
for t in 256 512; do
  SDB_DX_THREADS=$t timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:dx_gather -c 1 --csv python tools/prof_step.py --steps 1 2>/dev/null | grep -E "dx_gather" | awk -F'","' -v t=$t '{print t, $(NF-2), $(NF)}'
done
