#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture: DRAM bytes per launch of the captured kernels.

  ncu -i capture.ncu-rep --page raw --csv > raw.csv
  python tools/ncu_traffic.py raw.csv C4 threshold:k_threshold_tma [scale] [source-note]

bench.py quotes the entry of its dominant kernel as roofline.traffic (and null when there is none).  `scale` multiplies
the captured bytes when the capture ran a smaller batch than the config (per-launch traffic of these kernels is
proportional to the number of frames)."""
import csv
import json
import os
import sys


def main():
    raw, cfg, spec = sys.argv[1], sys.argv[2], sys.argv[3]
    scale = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    note = sys.argv[5] if len(sys.argv) > 5 else os.path.basename(raw)
    bench_name, kern = spec.split(":")
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n, dur = 0.0, 0, 0.0
    for r in rows[2:]:
        if kern not in r[col["Kernel Name"]]:
            continue
        rd = float(r[col["dram__bytes_read.sum"]]) * to_bytes[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * to_bytes[units[col["dram__bytes_write.sum"]]]
        tot += rd + wr
        dur += float(r[col["gpu__time_duration.sum"]])
        n += 1
    if not n:
        raise SystemExit("kernel %s not in %s" % (kern, raw))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    d = json.load(open(out)) if os.path.exists(out) else {}
    d.setdefault(cfg, {})[bench_name] = {"dram_bytes_per_launch": tot / n * scale, "kernel": kern, "launches_captured": n,
                                          "scaled_by": scale, "source": note}
    json.dump(d, open(out, "w"), indent=1, sort_keys=True)
    print(cfg, bench_name, d[cfg][bench_name])


if __name__ == "__main__":
    main()
