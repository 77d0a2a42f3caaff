"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (profiles/)."""
import collections, csv, sys

def main():
    src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
    skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    n_rows = 0
    for r in data:
        if len(r) <= vi:
            continue
        n_rows += 1
        if n_rows <= skip:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` launch list (cold-cache, serialised: compare shares, not absolutes).\n")
        f.write(f"Launches summarised: {n_rows - skip} (first {skip} skipped as warm-up).\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {t/1e3:.1f} | {t/1e3/n:.2f} | {100*t/tot:.1f}% |\n")
        f.write(f"\nTotal GPU time: {tot/1e6:.3f} ms\n")
    print(open(dst).read())

main()
