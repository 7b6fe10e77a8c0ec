"""Static SASS instruction count per source line (code-size view) from an ncu report.
    python profiles/ncu_static.py report.ncu-rep kernel-substring [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; want = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fn = fpath = None; cur = None
cnt = collections.Counter(); text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": continue
    if want not in (fn or ""): continue
    if r[2] == "-":
        cur = (fpath, r[0]); text[cur] = r[1].strip()[:100]
    elif cur is not None:
        cnt[cur] += 1
tot = sum(cnt.values()); print("static SASS instructions:", tot)
byfile = collections.Counter()
for (f, l), c in cnt.items(): byfile[f] += c
print(dict(byfile))
for (f, l), c in cnt.most_common(top):
    print(f"{f:>28s}:{l:<5s} {c:5d} | {text[(f,l)]}")
