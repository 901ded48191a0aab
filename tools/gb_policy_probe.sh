#!/bin/bash
# DRAM bytes of the dense accumulate kernel (G = 65536, L2-resident table) under L2 policy variants
for v in "none:0" "last:0" "normal:0" "unchanged:0" "none:2"; do
  pol=${v%%:*}; per=${v##*:}
  echo "== policy=$pol persist=$per"
  MB200_GB_POLICY=$pol MB200_GB_PERSIST=$per timeout 200 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gb_accumulate_tma_kernel -s 3 -c 1 python tools/gb_dense_probe.py 27 ${1:-65536} dense_nosmem 2>&1 | grep -E "dram__bytes|gpu__time|hit_rate"
done
