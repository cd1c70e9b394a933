#!/bin/bash
# One-box A/B of K8 v5 builds (tools/scratch/ab/lib*.so): short bench line + random-descriptor probe per build, parity tests on the candidates.
cd /root/repo
mkdir -p gpurun_out
LOG=gpurun_out/k8_sparse_ab.log
: > $LOG
cp slideo_b200/libslideo_b200.so /tmp/lib_keep.so
for f in A_head C2 B2 C1; do
  cp tools/scratch/ab/lib$f.so slideo_b200/libslideo_b200.so
  echo "== $f" | tee -a $LOG
  timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/ab_err_$f.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('frames/s', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'K8 Gpair/s', round(r['achieved'], 1), 'launch ms', round(r['avg_launch_ms'], 3), 'detect ms', round(d['detail']['ms_detect_per_step'], 1), 'knn ms', round(d['detail']['ms_knn_per_step'], 1), 'truth', d['detail']['frames_with_truth_match'])
" 2>&1 | tee -a $LOG
  timeout 60 python tools/prof_knn.py 303104 103000 4 2>&1 | tail -n 1 | tee -a $LOG
done
for f in C2 C1; do
  cp tools/scratch/ab/lib$f.so slideo_b200/libslideo_b200.so
  echo "== tests $f" | tee -a $LOG
  timeout 240 python -m pytest tests/test_gpu_knn.py tests/test_gpu_stream.py tests/test_gpu_fixtures.py -m gpu -x -q 2>&1 | tail -n 6 | tee -a $LOG
  if [ "$f" = "C2" ] && tail -n 3 $LOG | grep -q "passed" && ! tail -n 3 $LOG | grep -q "failed"; then break; fi
done
cp /tmp/lib_keep.so slideo_b200/libslideo_b200.so
