"""Per-instruction hot spots of an .ncu-rep captured with --import-source on (SASS view).
usage: python tools/ncu_source.py rep [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ins = []
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    try:
        ins.append((n, r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])))
    except ValueError:
        pass
tot_i = sum(i[2] for i in ins); tot_s = sum(i[3] for i in ins)
print("instructions", len(ins), "warp-instr executed", tot_i, "samples", tot_s)
ops = collections.Counter(); smp = collections.Counter()
for n, s, e, sa in ins:
    t = s.split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "")
    ops[op.split(".")[0]] += e; smp[op.split(".")[0]] += sa
print("by opcode (executed share | sample share):")
for k, v in ops.most_common(18):
    print(f"  {k:10s} {v/tot_i:6.3f} {smp[k]/max(tot_s,1):6.3f}")
# contiguous regions of 40 instructions with their executed share
print("regions (start idx: exec share, sample share, first instr):")
W = 40
for a in range(0, len(ins), W):
    seg = ins[a:a + W]
    print(f"  {a:5d}: {sum(i[2] for i in seg)/tot_i:6.3f} {sum(i[3] for i in seg)/max(tot_s,1):6.3f}  {seg[0][1][:60]}")
print("top sampled instructions:")
for n, s, e, sa in sorted(ins, key=lambda x: -x[3])[:top]:
    print(f"  {n:5d} {sa/max(tot_s,1):6.3f} {e/tot_i:6.3f} {s[:90]}")
