"""Per-kernel-family DRAM traffic and time from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
log of bench.py: writes profiles/traffic.json (read by bench.py for roofline.traffic) and prints a per-family table.

  python tools/ncu_traffic.py gpurun_out/<tag>_traffic.csv <source-label>
"""
import collections
import csv
import json
import os
import re
import sys

FAMILIES = [('conv3x3', r'conv3x3_kernel|thin_conv_kernel'), ('wgrad3x3', r'wgrad3x3_(tma_)?kernel|thin_wgrad_kernel'), ('bn_bwd', r'bn_bwd_(flat_|pool_)?kernel'),
            ('decoder_head', r'decoder_head_kernel'), ('gemm', r'gemm_kernel'),
            ('latent_fwd', r'latent_fwd_kernel'), ('latent_bwd', r'latent_bwd_kernel'), ('linear_f32', r'linear_f32_kernel')]


def to_bytes(v, unit):
    m = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    return v * m[unit]


def to_ms(v, unit):
    return v * {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}[unit]


def main(path, label):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ik, im, iu, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value'), hdr.index('ID')
    per = collections.defaultdict(dict)
    name = {}
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(',', ''))
        name[r[iid]] = r[ik]
        if r[im].startswith('dram__bytes'):
            per[r[iid]][r[im]] = to_bytes(v, r[iu])
        elif r[im].startswith('gpu__time'):
            per[r[iid]]['ms'] = to_ms(v, r[iu])
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for k, d in per.items():
        for f, pat in FAMILIES:
            if re.search(pat, name[k]) and 'pack' not in name[k]:
                fam[f][0] += 1
                fam[f][1] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
                fam[f][2] += d.get('ms', 0.0)
                break
    out = {}
    print(f'{"family":12s} {"launches":>8s} {"DRAM MB/launch":>15s} {"ms/launch":>10s} {"GB/s":>8s}')
    for f, (n, by, ms) in sorted(fam.items(), key=lambda kv: -kv[1][2]):
        out[f] = dict(dram_bytes_per_launch=round(by / n), launches_captured=n, ms_per_launch_under_ncu=round(ms / n, 4), source=label)
        print(f'{f:12s} {n:8d} {by / n / 1e6:15.1f} {ms / n:10.4f} {by / ms / 1e6 if ms else 0:8.1f}')
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'traffic.json')
    json.dump(out, open(dst, 'w'), indent=1)
    print('wrote', dst)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else os.path.basename(sys.argv[1]))
