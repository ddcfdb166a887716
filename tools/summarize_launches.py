"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`
launch list: per kernel count, time, time per launch, DRAM bytes per launch and DRAM rate."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui, mi = (hdr.index(x) for x in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Metric Name'))
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki][:56], {'n': 0, 't': 0.0, 'rd': 0.0, 'wr': 0.0})
        v, u, m = float(r[vi].replace(',', '')), r[ui], r[mi]
        if m == 'gpu__time_duration.sum':
            a['n'] += 1
            a['t'] += v / 1e6 if u == 'ns' else (v / 1e3 if u == 'us' else v)
        elif m == 'dram__bytes_read.sum':
            a['rd'] += v * scale[u]
        elif m == 'dram__bytes_write.sum':
            a['wr'] += v * scale[u]
    tot = sum(a['t'] for a in agg.values())
    print(f"total {tot:.2f} ms over {sum(a['n'] for a in agg.values())} launches (ncu serialised, cold cache: compare SHARES)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1]['t']):
        n = max(a['n'], 1)
        dram = f"  rd {a['rd'] / n / 1e6:8.1f} MB wr {a['wr'] / n / 1e6:8.1f} MB -> {(a['rd'] + a['wr']) / max(a['t'], 1e-9) / 1e9:5.2f} TB/s" if a['rd'] + a['wr'] > 0 else ""
        print(f"{a['t']:9.2f} ms {100 * a['t'] / tot:5.1f}%  x{a['n']:4d} {a['t'] / n * 1000:9.1f} us/launch{dram}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
