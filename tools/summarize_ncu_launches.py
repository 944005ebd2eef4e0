"""ncu `--csv` launch list (gpu__time_duration.sum [+ dram__bytes_read.sum, dram__bytes_write.sum]) -> per-kernel JSON.

usage: python tools/summarize_ncu_launches.py gpurun_out/launches_dram.csv profiles/r01_runN_dram_traffic_unetpp.json
The JSON is what bench.py reads for `roofline.traffic` (DRAM bytes per launch of the dominant kernel family)."""
import collections
import csv
import json
import re
import sys


def main(src: str, dst: str) -> None:
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ix = {n: h.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches: dict = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        name = re.sub(r"<.*|\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("gdl::", "")
        e = launches.setdefault(r[ix["ID"]], {"name": name, "us": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u, m = r[ix["Metric Unit"]], r[ix["Metric Name"]]
        if "time" in m:
            e["us"] = v / 1e3 if u.startswith("n") else (v if u.startswith("u") else v * 1e3)
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            e["rd" if "read" in m else "wr"] = v * mult
    agg: dict = collections.OrderedDict()
    for e in launches.values():
        a = agg.setdefault(e["name"], {"launches": 0, "ms": 0.0, "dram_read_GB": 0.0, "dram_write_GB": 0.0})
        a["launches"] += 1
        a["ms"] += e["us"] / 1e3
        a["dram_read_GB"] += e["rd"] / 1e9
        a["dram_write_GB"] += e["wr"] / 1e9
    total = sum(a["ms"] for a in agg.values())
    for a in agg.values():
        a["share"] = a["ms"] / total
        a["dram_GBps"] = (a["dram_read_GB"] + a["dram_write_GB"]) / a["ms"] * 1e3 if a["ms"] else 0.0
        a["dram_bytes_per_launch"] = (a["dram_read_GB"] + a["dram_write_GB"]) * 1e9 / a["launches"]
    out = {"source": src, "note": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
           "--clock-control none over one eager training step (cold-cache, serialised launches: shares, not absolutes)",
           "captured_launches": len(launches), "total_ms": total,
           "kernels": dict(sorted(agg.items(), key=lambda kv: -kv[1]["ms"]))}
    json.dump(out, open(dst, "w"), indent=1)
    for k, a in list(out["kernels"].items())[:12]:
        print(f"{k[:34]:34s} n={a['launches']:4d} {a['ms']:8.2f} ms {100 * a['share']:5.1f}%  {a['dram_GBps']:7.0f} GB/s")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
