"""Aggregate a train-step ncu launch list (tools/gpu_train_lists.sh) by kernel: last step only.
usage: train_list_summary.py list.csv [rows] [--last-step out.csv]   (--last-step: also write the raw CSV lines of the last step, for profiles/)"""
import collections, csv, sys
args = sys.argv[1:]
out_csv = None
if "--last-step" in args:
    i = args.index("--last-step")
    out_csv = args[i + 1]
    del args[i:i + 2]
lines = [l for l in open(args[0]) if not l.startswith("==")]
rows = [r for r in csv.reader(lines) if len(r) >= 15 and r[0].isdigit()]
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r[0], {"id": r[0], "k": r[4].split('(')[0].replace('void ', '').replace('hk::', '')})[r[12]] = float(r[14].replace(',', ''))
L = list(by.values())
idx = [i for i, l in enumerate(L) if l['k'] == 'stem_pack_kernel']
step = L[idx[-1]:]
if out_csv:
    keep = {l["id"] for l in step}
    with open(out_csv, "w") as f:
        f.write(lines[0])
        for l in lines[1:]:
            if l.split(",", 1)[0].strip('"') in keep:
                f.write(l)
agg = {}
for l in step:
    a = agg.setdefault(l['k'], [0, 0, 0]); a[0] += l['gpu__time_duration.sum']; a[1] += 1
    a[2] += l.get('dram__bytes_read.sum', 0) + l.get('dram__bytes_write.sum', 0)
tot = sum(a[0] for a in agg.values())
print("step ms", round(tot / 1e6, 3), "launches", len(step))
for k, (t, n, b) in sorted(agg.items(), key=lambda x: -x[1][0])[: int(args[1]) if len(args) > 1 else 40]:
    print(f"{k[:48]:48s} {n:4d} {t/1e6:8.3f} ms {100*t/tot:5.1f}%  avg {t/n/1e3:7.1f} us  {b/1e9:7.2f} GB {b/t if t else 0:6.0f} GB/s")
