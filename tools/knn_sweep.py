"""Developer probe: K8 throughput for several (nq, nt) with QR forced via SLIDEO_KNN_QR."""
import os, sys, json, subprocess
cfgs = [(131072, 103000), (303104, 103000), (606208, 103000), (65536, 1000000)]
for qr in ("4", "8"):
    for nq, nt in cfgs:
        env = dict(os.environ, SLIDEO_KNN_QR=qr)
        out = subprocess.run([sys.executable, "tools/prof_knn.py", str(nq), str(nt), "4"], env=env, capture_output=True, text=True)
        print("QR", qr, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
