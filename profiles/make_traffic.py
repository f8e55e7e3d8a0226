"""profiles/traffic.json from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log of
profiles/prof_driver.py: measured DRAM bytes per launch, summed per stage of ONE assembly (the last one in the log)."""
import csv, json, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn, mn, mv, idc = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    d = launch.setdefault(r[idc], {'name': r[kn]})
    d[r[mn]] = float(r[mv].replace(',', ''))
L = list(launch.values())
# one assembly = from a geometry kernel to the launch before the next geometry kernel
# round 2: the geometry kernel is fused into the first sweep (k_geo_sweep / its NVRTC build gsb_jit_geo): that launch opens an assembly
fused = any('k_geo_sweep' in d['name'] for d in L)
def opens(d): return 'k_geometry' in d['name'] or 'gsb_jit_geo' in d['name'] or 'k_geo_sweep' in d['name']
starts = [i for i, d in enumerate(L) if opens(d)]
seq = L[starts[-1]:]
seq = [d for d in seq if not any(t in d['name'] for t in ('k_spmv', 'k_cg_', 'k_diag', 'k_col_extent', 'k_spmv_classify'))]
def stage(d):
    n = d['name']
    if fused and opens(d): return 'sweep0'
    if 'k_geometry' in n or 'gsb_jit_geo' in n: return 'geometry'
    if 'k_vsweep' in n: return 'rhs'
    if 'TLast' in n or ', true' in n or ', 1, ' in n and 'TMass' in n: return 'sweep_last'
    if 'S1' in n: return 'sweep0'
    if 'S2' in n: return 'sweep1'
    return 'other'
out = collections.OrderedDict(); per = []
for d in seq:
    st = stage(d)
    b = d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    o = out.setdefault(st, {'dram_bytes': 0.0, 'ms': 0.0, 'launches': 0})
    o['dram_bytes'] += b; o['ms'] += d.get('gpu__time_duration.sum', 0) / 1e6; o['launches'] += 1
    per.append({'kernel': d['name'][:90], 'dram_read': d.get('dram__bytes_read.sum', 0), 'dram_write': d.get('dram__bytes_write.sum', 0), 'ms': d.get('gpu__time_duration.sum', 0) / 1e6})
res = {'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum (--clock-control none), one assembly of config 2; bytes and ms per stage (all launches of the stage)',
       'geometry': out.get('geometry', {}).get('dram_bytes'), 'sweep0': out.get('sweep0', {}).get('dram_bytes'), 'sweep1': out.get('sweep1', {}).get('dram_bytes'),
       'sweep2': out.get('sweep_last', {}).get('dram_bytes'), 'rhs': out.get('rhs', {}).get('dram_bytes'), 'stages': out, 'launches': per}
json.dump(res, open(sys.argv[2], 'w'), indent=1)
print(json.dumps(out, indent=1))
