"""Aggregates `ncu -i <rep> --page source --csv` of lattice_conv2_kernel by warp role (the code between the
USETMAXREG markers), stall reason and opcode.   python tools/ncu_source_summary.py source_page.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, E = ix["Warp Stall Sampling (All Samples)"], ix["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
inc = next(i for i, r in enumerate(data) if "USETMAXREG.TRY_ALLOC" in r[1])
dec = [i for i, r in enumerate(data) if "USETMAXREG.DEALLOC" in r[1]]
regions = {"compute warps": (inc, dec[0]), "UMMA issuer": (dec[0], dec[1]), "gather warps": (dec[1], len(data))}
total = sum(int(r[S]) for r in data)
print(rows[0][1] if len(rows[0]) > 1 else "", "\ninstructions", len(data), "samples", total)


def opcode(src):
    t = src.strip().split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


for name, (a, b) in regions.items():
    part = data[a:b]
    n = sum(int(r[S]) for r in part)
    print(f"\n== {name}: {n} samples ({n / total * 100:.1f} %), {b - a} instructions")
    agg = sorted(((sum(int(r[ix[s]]) for r in part), s) for s in stalls), reverse=True)
    print("   stall reasons: " + ", ".join(f"{s[6:]} {v / max(n, 1) * 100:.1f} %" for v, s in agg[:7]))
    ex = collections.Counter()
    for r in part:
        ex[opcode(r[1])] += int(r[E])
    te = sum(ex.values())
    print("   issued mix:    " + ", ".join(f"{k} {v / max(te, 1) * 100:.1f} %" for k, v in ex.most_common(12)))
    if name == "compute warps":
        ld = [i for i, r in enumerate(part) if "LDTM" in r[1]]
        if ld:
            lo, hi = max(ld[0] - 40, 0), ld[0] + 130
            e = sum(int(r[S]) for r in part[lo:hi])
            print(f"   accumulator drain (instructions {a + lo}..{a + hi}): {e / n * 100:.1f} % of this role's samples")
        for key in ("FENCE", "SYNCS", "NANOSLEEP"):
            v = sum(int(r[S]) for r in part if opcode(r[1]) == key)
            print(f"   samples on {key}: {v / n * 100:.1f} %")
