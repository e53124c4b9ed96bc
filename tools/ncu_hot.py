"""Hot source lines of one kernel in an .ncu-rep: instructions executed and stall samples per CUDA-C line."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
# The combined view lists a CUDA line followed by its SASS; fall back to sass-only aggregated by nothing if unsupported
rows = list(csv.reader(io.StringIO(out)))
hdr = None
agg = collections.OrderedDict()
cur = None
first_kernel_done = False
for r in rows:
    if len(r) > 3 and r[0] == "Address" or (len(r) > 3 and r[0] in ("Line No", "#")):
        if hdr is not None: first_kernel_done = True
        hdr = r; continue
    if hdr is None or len(r) != len(hdr) or first_kernel_done: continue
    d = dict(zip(hdr, r))
    src = d.get("Source", "")
    try:
        inst = int(d.get("Instructions Executed", "0") or 0); samp = int(d.get("# Samples", "0") or 0)
        thr = int(d.get("Thread Instructions Executed", "0") or 0)
    except ValueError:
        continue
    key = src
    a = agg.setdefault(key, [0, 0, 0]); a[0] += inst; a[1] += samp; a[2] += thr
tot_i = sum(a[0] for a in agg.values()) or 1; tot_s = sum(a[1] for a in agg.values()) or 1
print(f"total inst {tot_i} samples {tot_s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{a[1]*100/tot_s:5.1f}% smp {a[0]*100/tot_i:5.1f}% inst  thr/inst {a[2]/max(a[0],1):5.1f}  {k.strip()[:110]}")
