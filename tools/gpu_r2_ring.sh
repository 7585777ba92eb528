#!/bin/bash
# v3 walker: ring size (ranks in flight) against frame time
cd /root/repo
for ring in 128 256 512; do
  touch structure-slam-pointline_b200/csrc/line.cu
  make -C structure-slam-pointline_b200/csrc LINE_DEFS="-DSSLPL_V3_RING=$ring" 2>&1 | grep -E " error"
  for w in 8 16; do
    echo "ring $ring warps $w: $(timeout -s KILL 120 python tools/v3_stats.py $w 2>&1 | head -1 | grep -o "'lsd_regions': [0-9.]*")  1280: $(timeout -s KILL 200 python tools/v3_stats.py $w 1280 960 2>&1 | head -1 | grep -o "'lsd_regions': [0-9.]*")"
  done
done
