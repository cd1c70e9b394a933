#!/bin/bash
# A/B of builds of the library on the same GPU box: every tools/scratch/ab/lib*.so is swapped in in turn, two rounds.
# usage: bash tools/scratch/ab.sh [k10_probe shapes ...]      or      AB_CMD='python bench.py ...' bash tools/scratch/ab.sh
cd /root/repo
cp slideo_b200/libslideo_b200.so /tmp/lib_keep.so
for round in 1 2; do
  for f in tools/scratch/ab/lib*.so; do
    cp "$f" slideo_b200/libslideo_b200.so
    echo "== $(basename $f) (round $round)"
    if [ -n "$AB_CMD" ]; then timeout 300 bash -c "$AB_CMD" 2>&1 | tail -n 3 | cut -c1-400
    else timeout 40 python tools/k10_probe.py "$@" 2>&1 | cut -c1-110; fi
  done
done
cp /tmp/lib_keep.so slideo_b200/libslideo_b200.so
