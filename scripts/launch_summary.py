"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel: python scripts/launch_summary.py launches.csv"""
import csv, re, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
tot = 0.0
for r in rows:
    name = r[4].replace("void ", "").replace("at::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    m = re.search(r"(\w+Functor|\w+_kernel_cuda|\w+Kernel\w*|\w+_kernel)\b", name)
    base = re.sub(r"[<(].*", "", name)
    name = (base if base.startswith(("mobgt", "nvjet", "cutlass", "cusparse", "cublas")) or not m else base + ":" + re.search(r"(\w*Functor\w*|\w+_kernel\w*|\w+Kernel\w*|Op<[\w:]+)", name[len(base):] or name).group(1) if re.search(r"(\w*Functor\w*|\w+_kernel\w*|\w+Kernel\w*|Op<[\w:]+)", name[len(base):] or name) else base)[:70]
    ns = float(r[14])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
    tot += ns
ours = sum(v[1] for k, v in agg.items() if k.startswith("mobgt::") or k.startswith("k"))
print(f"# {len(rows)} launches, {tot/1e6:.2f} ms of kernel time (cold-cache, serialised under ncu); libmobgt share {100*ours/tot:.1f}%")
print(f"{'kernel':72s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>6s}")
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {c:8d} {ns/1e3:10.1f} {ns/1e3/c:9.1f} {100*ns/tot:5.1f}%")
