"""Per-source-line instruction and stall-sample totals from an ncu report (cuda,sass view).
    python profiles/ncu_lines.py report.ncu-rep [kernel-substring] [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fn = None; fpath = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])  # (file,line) -> inst, samples, thread inst, text
tot = collections.defaultdict(lambda: [0, 0])
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; iL = 0; iS = 1; iN = hdr.index("# Samples"); iE = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or want not in (fn or ""): continue
    if r[2] != "-": continue  # SASS rows repeat the totals of the source line row
    try:
        e = int(r[iE].replace(",", "")); s = int(r[iN].replace(",", "")); t = int(r[iT].replace(",", ""))
    except ValueError:
        continue
    key = (fn.split("(")[0][-40:], fpath, int(r[iL]))
    a = agg[key]; a[0] += e; a[1] += s; a[2] += t; a[3] = r[iS].strip()[:110]
    tot[key[0]][0] += e; tot[key[0]][1] += s
for k, (e, s) in tot.items():
    print(f"== {k}: warp-inst {e:,} samples {s:,}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    te, ts = tot[key[0]]
    print(f"{key[1]:>16s}:{key[2]:<4d} inst {a[0]/te*100:5.1f}% samp {a[1]/max(ts,1)*100:5.1f}% thr/inst {a[2]/max(a[0],1):4.1f} | {a[3]}")

# coarse buckets: file -> [(first_line, last_line, label)] given as extra args "file:lo-hi=label"
buckets = [a for a in sys.argv[4:] if "=" in a]
if buckets:
    print("---- buckets")
    for b in buckets:
        spec, label = b.split("=")
        f, rng = spec.split(":"); lo, hi = (int(x) for x in rng.split("-"))
        e = sum(a[0] for k, a in agg.items() if k[1] == f and lo <= k[2] <= hi)
        s = sum(a[1] for k, a in agg.items() if k[1] == f and lo <= k[2] <= hi)
        te = sum(v[0] for v in tot.values()); ts = sum(v[1] for v in tot.values())
        print(f"{label:30s} inst {e/te*100:5.1f}%  samples {s/max(ts,1)*100:5.1f}%")
