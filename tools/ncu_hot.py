"""Top stall locations of an `ncu --page source --csv` dump (optionally gzipped): python tools/ncu_hot.py file.csv[.gz] [N]"""
import csv, gzip, io, sys
path = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hdr = rows[1]
ia, isrc, iall, iexe = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > iexe]
tot = sum(int(r[iall] or 0) for r in body)
print(f"{len(body)} instructions, {tot} samples, {sum(int(r[iexe] or 0) for r in body)} warp instructions executed")
idx = {id(r): i for i, r in enumerate(body)}
for r in sorted(body, key=lambda r: -int(r[iall] or 0))[:N]:
    i = idx[id(r)]
    print(f"{i:5d} {100.0 * int(r[iall] or 0) / tot:5.1f}%  exec {int(r[iexe] or 0):8d}  {r[isrc].strip()[:100]}")

# stall reasons summed over the instructions executed by the hot warps (exec count >= --min-exec of the maximum)
reasons = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
if reasons:
    mx = max(int(r[iexe] or 0) for r in body)
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.2
    hot = [r for r in body if int(r[iexe] or 0) >= thr * mx]
    tot_hot = sum(int(r[iall] or 0) for r in hot)
    print(f"\nstall reasons over {len(hot)} instructions executed >= {thr:.2f} x max ({100.0 * tot_hot / tot:.1f}% of all samples):")
    agg = {h: sum(int(r[hdr.index(h)] or 0) for r in hot) for h in reasons}
    for h, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        if v:
            print(f"   {h:28s} {100.0 * v / max(tot_hot, 1):5.1f}%")
