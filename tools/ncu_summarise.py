"""Summaries of the ncu CSV logs a GPU session brings back (run in the build container):
  python tools/ncu_summarise.py launches gpurun_out/l_launch_list.csv "<command>"  > profiles/r2_launch_list_summary.txt
  python tools/ncu_summarise.py traffic  gpurun_out/l_traffic.csv "<command>"      > profiles/r2_traffic.json
launches: per-kernel launch count, total / share / average of gpu__time_duration.sum.
traffic : per-kernel mean dram__bytes_read.sum + dram__bytes_write.sum per launch (orth / spmv keys,
          read by bench.py for `roofline.traffic`)."""
import csv
import json
import re
import sys


def rows(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name[:74]


def main():
    mode, path, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rs = rows(path)
    if mode == "launches":
        agg = {}
        for r in rs:
            if r["Metric Name"] != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            v = v / 1e3 if r["Metric Unit"] == "ns" else (v if r["Metric Unit"] in ("us", "usecond") else v * 1e3)
            a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        print("# " + cmd)
        print("# (cold-cache, serialised per-launch times: compare SHARES with bench.py's live CUDA-event shares)")
        print("%-76s %5s %12s %7s %10s" % ("kernel", "n", "total_us", "share", "avg_us"))
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("%-76s %5d %12.1f %6.1f%% %10.1f" % (k, n, us, 100 * us / tot, us / n))
    else:
        per = {}
        for r in rs:
            key = "orth" if "orth_kernel" in r["Kernel Name"] else ("spmv" if "spmv_staged" in r["Kernel Name"] else None)
            if key is None:
                key = short(r["Kernel Name"]).replace("void ", "")
            d = per.setdefault(key, {})
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            if r["Metric Name"].startswith("dram__bytes"):
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            elif r["Metric Name"].startswith("gpu__time"):
                v *= {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1)
            d.setdefault(r["Metric Name"], []).append(v)
        out = {}
        for key, d in per.items():
            rd, wr = d.get("dram__bytes_read.sum", []), d.get("dram__bytes_write.sum", [])
            n = max(len(rd), 1)
            out[key] = {"launches": len(rd), "dram_bytes_read_per_launch": sum(rd) / n,
                        "dram_bytes_write_per_launch": sum(wr) / n, "traffic_per_launch": (sum(rd) + sum(wr)) / n,
                        "avg_us_under_ncu": sum(d.get("gpu__time_duration.sum", [0])) / n}
        out["command"] = cmd
        print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
