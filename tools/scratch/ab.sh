#!/bin/bash
# A/B of two builds of the library on the same GPU box: tools/scratch/ab/lib{A,B}.so are swapped in turn
cd /root/repo
cp slideo_b200/libslideo_b200.so /tmp/lib_keep.so
for round in 1 2; do
  for v in A B; do
    cp tools/scratch/ab/lib$v.so slideo_b200/libslideo_b200.so
    echo "== $v (round $round)"
    timeout 40 python tools/k10_probe.py "$@" 2>&1 | cut -c1-110
  done
done
cp /tmp/lib_keep.so slideo_b200/libslideo_b200.so
