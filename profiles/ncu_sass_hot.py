"""Per-opcode and per-instruction breakdown of `ncu --page source --csv --print-source sass` output."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]
ie, isamp, isrc = h.index('Instructions Executed'), h.index('# Samples'), h.index('Source')
ilong = h.index('stall_long_sb'); inoi = h.index('stall_no_inst'); iwait = h.index('stall_wait'); ishort = h.index('stall_short_sb')
data = []
for k, r in enumerate(rows[hi + 1:]):
    try:
        data.append((int(r[ie]), int(r[isamp]), k, r[isrc].strip(), int(r[ilong]), int(r[inoi]), int(r[iwait]), int(r[ishort])))
    except Exception:
        pass
ti = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("static instructions", len(data), "warp instructions", ti, "samples", ts)
byop = collections.defaultdict(lambda: [0, 0])
for d in data:
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', d[3]); op = m.group(2) if m else '?'
    byop[op][0] += d[0]; byop[op][1] += d[1]
print("--- by opcode (inst%, samples%)")
for op, v in sorted(byop.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0]/ti*100:5.1f}% {v[1]/max(ts,1)*100:5.1f}%  {op}")
print("--- top stall instructions: samples% [long, noinst, wait, short] idx source")
for d in sorted(data, key=lambda d: -d[1])[:top]:
    print(f"{d[1]/max(ts,1)*100:5.1f}% {d[4:8]} #{d[2]} x{d[0]} {d[3][:90]}")
