"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: last forward+backward sequence."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
i = next(k for k, r in enumerate(rows) if "Kernel Name" in r)
h = rows[i]
seq = [(r[h.index("Kernel Name")].replace("void ", "").replace("mcacq::", "").split("(")[0][:48],
        float(r[h.index("Metric Value")].replace(",", "")) / 1e6) for r in rows[i + 2:] if len(r) > h.index("Metric Value")]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
tot = sum(v for _, v in seq[-n:])
for name, v in seq[-n:]:
    print(f"{v:9.3f} ms {100*v/tot:5.1f}%  {name}")
print(f"{tot:9.3f} ms total of last {n} launches")
