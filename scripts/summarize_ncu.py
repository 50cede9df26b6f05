"""Turn ncu output into the markdown tables kept under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
  python scripts/summarize_ncu.py full gpurun_out/x.ncu-rep         > profiles/rNN_x.md

`launches` reads the `--metrics gpu__time_duration.sum --csv` log; `full` shells out to
`ncu -i <rep> --page raw --csv` and prints the speed-of-light lines that bench.py's roofline quotes.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("<unnamed>::", "")
    name = re.sub(r"\(.*$", "", name)
    return name[:70]


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("ns", "nsecond"):
            v /= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e3
        rows.append((short(r["Kernel Name"]), v))
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in rows)
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f}% |")
    print(f"\nTotal: {tot / 1e3:.2f} ms over {len(rows)} launches.")


WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_executed_pipe_tensor_subunit_tc.sum", "tc_inst"),
    ("sm__pipe_tensor_subunit_tc_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__cycles_active.avg", "cycles"),
]


def full(path):
    """`path`: a .ncu-rep, or the `ncu -i <rep> --page raw --csv` dump of one (scripts/profile_round.sh reduces the
    captures to CSV on the GPU box: the reports themselves are too large to bring back)."""
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    while rows and (not rows[0] or rows[0][0] != "ID"):
        rows.pop(0)
    hdr, units = rows[0], rows[1]
    cols = [(lab, hdr.index(m)) for m, lab in WANT if m in hdr]
    ki = hdr.index("Kernel Name")
    print("| # | kernel | " + " | ".join(f"{lab} ({units[i]})" if units[i] else lab for lab, i in cols) + " |")
    print("|---|---|" + "---:|" * len(cols))
    for n, r in enumerate(rows[2:]):
        vals = []
        for _, i in cols:
            try:
                vals.append(f"{float(r[i].replace(',', '')):.4g}")
            except ValueError:
                vals.append(r[i])
        print(f"| {n} | `{short(r[ki])}` | " + " | ".join(vals) + " |")
    if "--tensor" in sys.argv:
        tens = [h for h in hdr if "tensor" in h or "tmem" in h.lower()]
        print("\nTensor / TMEM metrics present in the capture:")
        for h in tens[:24]:
            i = hdr.index(h)
            print(f"* `{h}` ({units[i]}): " + ", ".join(r[i] for r in rows[2:]))


def traffic(path):
    """Average DRAM bytes (read + write) per launch over the launches of a raw-CSV capture -> one JSON object."""
    import json
    out = open(path).read()
    rows = list(csv.reader(io.StringIO(out)))
    while rows and (not rows[0] or rows[0][0] != "ID"):
        rows.pop(0)
    hdr, units = rows[0], rows[1]
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for r in rows[2:]:
        tot += float(r[ir].replace(",", "")) * mul.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * mul.get(units[iw], 1.0)
    n = len(rows) - 2
    print(json.dumps(dict(dram_bytes_per_launch=tot / max(n, 1), launches=n, source=os.path.basename(path))))


if __name__ == "__main__":
    import os
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
