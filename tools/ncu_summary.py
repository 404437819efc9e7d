"""Turn raw ncu csv output (gpurun_out/, scratch) into the small summaries committed under profiles/.

    python tools/ncu_summary.py launches RAW.csv OUT.csv     # --metrics gpu__time_duration.sum launch list
    python tools/ncu_summary.py full RAW.csv OUT.csv [OUT.json]   # --set full --page raw --csv capture
    python tools/ncu_summary.py merge OUT_RAW.csv BASE_RAW.csv UPDATE_RAW.csv REGEX   # re-captured kernels replace their rows
    python tools/ncu_summary.py traffic OUT.json RAW.csv [RAW.csv ...]             # per-family DRAM bytes per launch

`launches`: one line per kernel (launch count, total / mean / min duration, share of the serialised kernel time).
`full`: one line per kernel instantiation (mean over its launches) with the metrics DESIGN.md / bench.py quote:
duration, DRAM bytes, cache hit rates, pipe utilisation, occupancy, issue activity and the warp-state stall reasons.
The optional json holds per-launch DRAM traffic (bench.py's `roofline.traffic`).
"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict

csv.field_size_limit(10 ** 9)

FULL_METRICS = [
    ('gpu__time_duration.sum', 'us'),
    ('launch__grid_size', ''),
    ('launch__block_size', ''),
    ('launch__registers_per_thread', ''),
    ('launch__shared_mem_per_block', 'KB'),
    ('dram__bytes_read.sum', 'MB'),
    ('dram__bytes_write.sum', 'MB'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', '%'),
    ('lts__t_sector_hit_rate.pct', '%'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', '%'),
    ('l1tex__t_sector_hit_rate.pct', '%'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', '%'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', '%'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', '%'),
    ('sm__issue_active.avg.pct_of_peak_sustained_elapsed', '%'),
    ('sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed', '%'),
    ('sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active', '%'),
    ('smsp__inst_executed.sum', ''),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', ''),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', ''),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', ''),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', ''),
    ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', ''),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', ''),
    ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', ''),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', ''),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', ''),
    ('smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio', ''),
]
SCALE = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}


def short(name):
    name = re.sub(r'^void ', '', name)
    if name.startswith('at::') or name.startswith('cudnn') or 'cutlass' in name:
        return name.split('<')[0].split('(')[0][:60]
    return re.sub(r'\(.*$', '', name).replace('srf::', '')


def rows_of(path):
    return list(csv.reader(l for l in open(path, errors='replace') if l.startswith('"')))


def launches(raw, out):
    rows = rows_of(raw)
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = OrderedDict()
    for r in rows[1:]:
        us = float(r[vi].replace(',', '')) * SCALE.get(r[ui], 1.0)
        a = agg.setdefault(short(r[ki]), [0, 0.0, 1e30])
        a[0] += 1
        a[1] += us
        a[2] = min(a[2], us)
    total = sum(a[1] for a in agg.values())
    with open(out, 'w') as f:
        f.write('kernel,launches,total_us,mean_us,min_us,share_of_serialised_kernel_time\n')
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{a[0]},{a[1]:.1f},{a[1] / a[0]:.2f},{a[2]:.2f},{a[1] / total:.4f}\n')
        f.write(f'"TOTAL",{sum(a[0] for a in agg.values())},{total:.1f},,,1.0\n')
    print(f'{out}: {len(agg)} kernels, {total / 1e3:.2f} ms serialised')


def full(raw, out, out_json=None):
    rows = rows_of(raw)
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    col = {}
    for m, _ in FULL_METRICS:
        if m in hdr:
            col[m] = hdr.index(m)
    agg = OrderedDict()
    for r in rows[2:]:
        key = short(r[ki]) + ' grid=' + r[hdr.index('Grid Size')].replace(' ', '')
        d = agg.setdefault(key, defaultdict(list))
        for m, want in FULL_METRICS:
            if m not in col or r[col[m]] == '':
                continue
            v = float(r[col[m]].replace(',', ''))
            u = units[col[m]]
            if want == 'us':
                v *= SCALE.get(u, 1.0)
            elif want == 'MB':
                v *= SCALE.get(u, 1.0)
            elif want == 'KB':
                v *= {'byte/block': 1e-3, 'Kbyte/block': 1.0}.get(u, 1.0)
            d[m].append(v)
    names = [m for m, _ in FULL_METRICS if m in col]
    with open(out, 'w') as f:
        f.write('kernel,launches,' + ','.join(f'{m}[{u}]' if u else m for m, u in FULL_METRICS if m in col) + '\n')
        for k, d in agg.items():
            n = len(d[names[0]])
            f.write(f'"{k}",{n},' + ','.join(f'{sum(d[m]) / max(1, len(d[m])):.4g}' for m in names) + '\n')
    if out_json:
        js = {'source': f'ncu --set full --clock-control none ({raw}); per-launch means; ncu flushes caches before every replay, so reads are cold-cache',
              'kernels': {k: {'launches': len(d['gpu__time_duration.sum']),
                              'dram_read_bytes': int(1e6 * sum(d['dram__bytes_read.sum']) / max(1, len(d['dram__bytes_read.sum']))),
                              'dram_write_bytes': int(1e6 * sum(d['dram__bytes_write.sum']) / max(1, len(d['dram__bytes_write.sum']))),
                              'ncu_duration_us': round(sum(d['gpu__time_duration.sum']) / len(d['gpu__time_duration.sum']), 2)}
                          for k, d in agg.items()}}
        json.dump(js, open(out_json, 'w'), indent=1)
    print(f'{out}: {len(agg)} kernel instantiations')


FAMILY_OF = [          # kernel-name regex -> bench.py kernel family (only the unambiguous ones)
    (r'img_roi_cl_kernel', 'image RoIAlign (6 cameras)'), (r'bev_roi_cl_kernel', 'BEV RoIAlign'),
    (r'dynconv_interact', 'DynamicConv interaction'), (r'mha_attention', 'self-attention core'),
    (r'igemm_umma_kernel<16, 32, 1', 'sparse conv 16->32'), (r'igemm_umma_kernel<32, 32, 1', 'sparse conv 32->32'),
    (r'igemm_umma_kernel<32, 64, 1', 'sparse conv 32->64'), (r'igemm_umma_kernel<64, 64, 1', 'sparse conv 64->64'),
    (r'igemm_umma_kernel<64, 128, 1', 'sparse conv 64->128'), (r'spconv16_warp_kernel', 'sparse conv 16->16'),
    (r'rulebook_(fast_)?kernel', 'rulebook build'),
]


def traffic(out_json, *raws):
    """profiles/r02_ncu_traffic.json: per-launch DRAM bytes of the kernel families bench.py reports, from ncu --set full captures."""
    fam = OrderedDict()
    for raw in raws:
        rows = rows_of(raw)
        hdr, units = rows[0], rows[1]
        ki = hdr.index('Kernel Name')
        cr, cw, ct = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
        for r in rows[2:]:
            name = next((f for pat, f in FAMILY_OF if re.search(pat, r[ki])), None)
            if name is None or r[cr] == '':
                continue
            d = fam.setdefault(name, dict(launches=0, rd=0.0, wr=0.0, us=0.0, sources=set()))
            d['launches'] += 1
            d['rd'] += float(r[cr].replace(',', '')) * SCALE.get(units[cr], 1.0) * 1e6
            d['wr'] += float(r[cw].replace(',', '')) * SCALE.get(units[cw], 1.0) * 1e6
            d['us'] += float(r[ct].replace(',', '')) * SCALE.get(units[ct], 1.0)
            d['sources'].add(raw)
    js = {'source': 'ncu --set full --clock-control none captures of bench.py / tools/bench_kernels.py launches; per-launch means; '
                    'ncu flushes caches before every replay, so reads are cold-cache',
          'families': {k: dict(launches=d['launches'], dram_read_bytes_per_launch=int(d['rd'] / d['launches']),
                               dram_write_bytes_per_launch=int(d['wr'] / d['launches']), ncu_duration_us=round(d['us'] / d['launches'], 2),
                               raw=sorted(d['sources'])) for k, d in fam.items()}}
    json.dump(js, open(out_json, 'w'), indent=1)
    print(f'{out_json}: {len(fam)} families')


def merge(out, base, update, pattern):
    """Replace the rows of `base` whose kernel name matches `pattern` by the rows of `update` (a later capture of the kernels that
    changed), matching columns by metric NAME (two ncu runs can differ by a column)."""
    a, b = rows_of(base), rows_of(update)
    ha, hb = a[0], b[0]
    ib = {n: i for i, n in enumerate(hb)}
    ki = ha.index('Kernel Name')
    pat = re.compile(pattern)
    rows = [ha, a[1]] + [r for r in a[2:] if not pat.search(r[ki])]
    rows += [[r[ib[n]] if n in ib else '' for n in ha] for r in b[2:]]
    with open(out, 'w', newline='') as f:
        csv.writer(f, quoting=csv.QUOTE_ALL).writerows(rows)
    print(f'{out}: {len(rows) - 2} launches')


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'traffic': traffic, 'merge': merge}[sys.argv[1]](*sys.argv[2:])
