"""profiles/traffic.json from ncu --set full captures: DRAM bytes (read + write) per launch, per bench.py conv class.

usage: python tools/ncu_traffic.py out.json class[,class..]=report.ncu-rep ...
The i-th kernel of a report belongs to the i-th listed class (one class name = every kernel of the report); a class
seen several times gets the mean.  Also prints a one-line summary per kernel (duration, dram bytes, sm / dram %).
"""
import csv
import json
import subprocess
import sys


def rows(rep):
    # a .ncu-rep, or the `ncu -i rep --page raw --csv` text saved on the GPU box (reports above the gpurun size cap stay there)
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    return hdr, units, r[2:]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    dst = sys.argv[1]
    acc = {}
    lines = []
    for spec in sys.argv[2:]:
        classes, rep = spec.split("=")
        classes = classes.split(",")
        hdr, units, data = rows(rep)
        def col(name):      # exact name, else the first column ending with it (ncu prefixes some with a unit path)
            return hdr.index(name) if name in hdr else next(i for i, h in enumerate(hdr) if h.endswith(name))
        ix = {k: col(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                  "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                                  "dram__throughput.avg.pct_of_peak_sustained_elapsed")}
        for i, row in enumerate(data):
            cls = classes[min(i, len(classes) - 1)]
            b = to_bytes(row[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
                to_bytes(row[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            acc.setdefault(cls, []).append(b)
            lines.append("%s #%d %-8s %s  time=%s %s  dram=%.1f MB  sm%%=%s dram%%=%s" % (
                rep.split("/")[-1], i, cls, row[ix["Kernel Name"]][:60], row[ix["gpu__time_duration.sum"]],
                units[ix["gpu__time_duration.sum"]], b / 1e6, row[ix["sm__throughput.avg.pct_of_peak_sustained_elapsed"]],
                row[ix["dram__throughput.avg.pct_of_peak_sustained_elapsed"]]))
    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches of the class) "
                       "from ncu --set full captures (profiles/*ncu_full*); keyed by bench.py conv class"}
    for k, v in acc.items():
        out[k] = sum(v) / len(v)
    json.dump(out, open(dst, "w"), indent=1)
    print("\n".join(lines))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
