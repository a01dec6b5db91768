"""Attribute the divergence-management instructions (BSSY / BSYNC / BRA / NOP) and the opcode mix of an ncu report to
source lines.   usage: python tools/ncu_divergence.py report.ncu-rep path/to/source.cu [top]"""
import collections, csv, io, re, subprocess, sys
rep, cu = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, T = None, 0
agg = collections.defaultdict(collections.Counter)
mix = collections.Counter()
for r in rows[3:]:
    if len(r) < 8:
        continue
    if r[0].strip().isdigit():
        cur = int(r[0]); continue
    if r[2].startswith("0x"):
        try:
            n = int(r[7])
        except ValueError:
            continue
        T += n
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3].strip())
        o = m.group(2) if m else "?"
        mix[o] += n
        if o in ("BSSY", "BSYNC", "BRA", "NOP"):
            agg[cur][o] += n
text = open(cu).read().split("\n")
print(f"# {rep}: {T} warp instructions")
print("# opcode mix:", ", ".join(f"{o} {100 * v / T:.2f}%" for o, v in mix.most_common(16)))
items = sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))
for l, c in items[:top]:
    print(f"L{l:4d} BSSY {100 * c['BSSY'] / T:.2f} BSYNC {100 * c['BSYNC'] / T:.2f} BRA {100 * c['BRA'] / T:.2f} NOP {100 * c['NOP'] / T:.2f} | {text[l - 1].strip()[:100]}")
