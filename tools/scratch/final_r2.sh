#!/bin/bash
# Last GPU call of the round: A/B of the entries-prefetch build, full GPU suite + smoke + default bench on the winner, ncu lists (best effort).
cd /root/repo
mkdir -p gpurun_out
LOG=gpurun_out/final_r2.log
: > $LOG
probe() {  # $1 = lib tag -> prints frames/s of a short bench line
  cp tools/scratch/ab/lib$1.so slideo_b200/libslideo_b200.so
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/final_err_$1.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
sys.stderr.write('frames/s %.1f e2e %.1f K8 Gpair/s %.1f launch ms %.3f knn ms %.1f truth %s walk %s\n' % (d['value'], d['e2e']['value'], r['achieved'], r['avg_launch_ms'], d['detail']['ms_knn_per_step'], d['detail']['frames_with_truth_match'], json.dumps(r.get('list_walk'))))
print('%.1f' % d['value'])
" 2>>$LOG
}
echo "== C1" >> $LOG; VC=$(probe C1); echo "== P1" >> $LOG; VP=$(probe P1)
echo "C1 $VC P1 $VP" | tee -a $LOG
WIN=C1
if [ -n "$VP" ] && [ -n "$VC" ] && python -c "import sys; sys.exit(0 if float('$VP') > 1.01 * float('$VC') else 1)"; then WIN=P1; fi
suite() {
  cp tools/scratch/ab/lib$1.so slideo_b200/libslideo_b200.so
  echo "== suite on $1" | tee -a $LOG
  timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8 | tee gpurun_out/r2_gpu_suite8.log | tee -a $LOG
  grep -q " passed" gpurun_out/r2_gpu_suite8.log && ! grep -q "failed\|error" gpurun_out/r2_gpu_suite8.log
}
if ! suite $WIN; then
  if [ $WIN = P1 ]; then WIN=C1; suite C1 || echo "SUITE FAILED ON C1 TOO" | tee -a $LOG; else echo "SUITE FAILED" | tee -a $LOG; fi
fi
echo "winner $WIN at $SECONDS s" | tee -a $LOG | tee gpurun_out/final_choice.txt
cp tools/scratch/ab/lib$WIN.so slideo_b200/libslideo_b200.so
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee -a $LOG
timeout 400 python bench.py > gpurun_out/r2_bench_r.json 2> gpurun_out/r2_bench_r.err; tail -c 600 gpurun_out/r2_bench_r.json | tee -a $LOG
# best effort from here on
[ $SECONDS -lt 300 ] && timeout 110 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 1 --warmup 3 --frames 192 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_final.log 2>&1
echo "ncu list rc $? at $SECONDS s" | tee -a $LOG
[ $SECONDS -lt 360 ] && timeout 75 ncu --set full --clock-control none --import-source on -k regex:knn5_kernel -s 3 -c 1 -o gpurun_out/r2_k8v5_sparse -f python tools/prof_frames.py 320 50 2 > gpurun_out/ncu_k8_sparse.log 2>&1
echo "ncu full rc $? at $SECONDS s" | tee -a $LOG
