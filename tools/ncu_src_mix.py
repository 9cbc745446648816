"""Instruction mix and stall samples per SASS opcode from `ncu -i X.ncu-rep --page source --csv`.
usage: python tools/ncu_src_mix.py src.csv [n_tets]"""
import csv, sys, collections
f = sys.argv[1]; n_tets = float(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(open(f)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
names = rows[hdr]
col = {n: i for i, n in enumerate(names)}
mix = collections.Counter(); samples = collections.Counter(); stall = collections.Counter()
stall_cols = [n for n in names if n.startswith("stall_") and "Not Issued" not in n]
tot_inst = 0
for r in rows[hdr + 1:]:
    if len(r) < len(names): continue
    src = r[col["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    base = op.split(".")[0]
    if base in ("MUFU", "F2F", "LDS", "STS", "LDG", "RED", "ATOMG", "LDGSTS", "BAR", "SYNCS", "SHFL", "I2F", "F2I", "LDSM"):
        key = ".".join(op.split(".")[:3]) if base in ("LDS", "STS", "MUFU", "F2F", "BAR", "SYNCS", "RED") else base
    else:
        key = base
    n = float(r[col["Instructions Executed"]] or 0)
    mix[key] += n; tot_inst += n
    samples[key] += float(r[col["# Samples"]] or 0)
    for s in stall_cols:
        stall[s] += float(r[col[s]] or 0)
tot_s = sum(samples.values())
print(f"total warp instructions {tot_inst:.4g}" + (f" = {tot_inst * 32 / n_tets:.0f} thread-instr per tet" if n_tets else ""))
print(f"{'opcode':28s} {'warp instr':>12s} {'/tet':>8s} {'% samples':>10s}")
for k, v in mix.most_common(45):
    print(f"{k:28s} {v:12.4g} {(v * 32 / n_tets if n_tets else 0):8.1f} {100 * samples[k] / max(tot_s, 1):10.1f}")
print("stall reasons (share of samples):")
ts = sum(stall.values())
for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:12]:
    print(f"  {k:28s} {100 * v / max(ts, 1):6.1f} %")
