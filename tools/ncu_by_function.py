"""Aggregate the stall samples / executed instructions of an ncu source page by the device function the source line
belongs to.   usage: python tools/ncu_by_function.py report.ncu-rep path/to/source.cu"""
import collections, csv, io, re, subprocess, sys
rep, cu = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
src = []
cur_file = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1]
    if len(r) > 7 and r[0].strip().isdigit() and cur_file and cur_file.endswith(cu.split("/")[-1]):
        try:
            src.append((int(r[0]), r[1].strip(), int(r[4] or 0), int(r[7] or 0)))
        except ValueError:
            pass
text = open(cu).read().split("\n")
funcs = []
for i, l in enumerate(text, 1):
    m = re.match(r"^(?:__device__|__global__|__host__|template).*?([a-zA-Z_0-9]+)\(", l)
    if m and not l.startswith(" ") and "template" not in l.split("(")[0]:
        funcs.append((i, m.group(1)))
def func_of(ln):
    name = "?"
    for i, n in funcs:
        if i <= ln: name = n
    return name
tot_s = sum(s[2] for s in src) or 1; tot_i = sum(s[3] for s in src) or 1
agg, aggi = collections.Counter(), collections.Counter()
for ln, t, s, i in src:
    agg[func_of(ln)] += s; aggi[func_of(ln)] += i
print(f"# {rep}: stall samples {tot_s}, warp instructions {tot_i}")
for k, v in agg.most_common(30):
    print(f"{k:26s} stall {100 * v / tot_s:5.1f} %   inst {100 * aggi[k] / tot_i:5.1f} %")
print("# top lines by stall samples")
for ln, t, s, i in sorted(src, key=lambda x: -x[2])[:30]:
    print(f"L{ln:<4d} stall {100 * s / tot_s:4.1f} %  inst {100 * i / tot_i:4.1f} %  {t[:110]}")
