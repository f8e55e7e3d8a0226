"""Top source lines by executed instructions / stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# sections start at header rows beginning with "Line No"; the first (source-level) table of the file is used
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
h = rows[starts[0]]
ie = h.index('Instructions Executed'); isamp = h.index('# Samples')
end = starts[1] if len(starts) > 1 else len(rows)
data = []
for r in rows[starts[0] + 1:end]:
    try:
        data.append((int(r[ie]), int(r[isamp]), r[0], r[1][:120]))
    except Exception:
        pass
ti = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("total warp instructions", ti, "samples", ts)
print("--- by instructions"); [print(f"{d[0]/ti*100:5.1f}% inst {d[1]/max(ts,1)*100:5.1f}% smp  L{d[2]:>4} {d[3]}") for d in sorted(data, reverse=True)[:top]]
print("--- by stall samples"); [print(f"{d[0]/ti*100:5.1f}% inst {d[1]/max(ts,1)*100:5.1f}% smp  L{d[2]:>4} {d[3]}") for d in sorted(data, key=lambda d: -d[1])[:top]]
