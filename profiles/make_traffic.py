"""Regenerate profiles/traffic.json (the `roofline.traffic` field of the bench line) from an ncu raw-page export:

    ncu --set full --clock-control none -k regex:count_reads_kernel -s <warm-up launches> -c 1 -o gpurun_out/prof_count python bench.py --config c2 ...
    ncu -i gpurun_out/prof_count.ncu-rep --page raw --csv > raw.csv
    python profiles/make_traffic.py raw.csv c2 [source note]

Takes dram__bytes_read.sum + dram__bytes_write.sum of the first count_reads_kernel launch in the export and stores it under
count_reads_kernel_dram_bytes_per_launch_<config>, next to where it came from.  Other keys of the file are kept."""
import csv
import json
import os
import sys

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    raw, config = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(raw)
    rows = list(csv.reader(open(raw)))
    head, units = rows[0], rows[1]
    name = head.index("Kernel Name")
    for row in rows[2:]:
        if "count_reads_kernel" in row[name]:
            total = 0.0
            parts = {}
            for metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = head.index(metric)
                parts[metric] = float(row[i].replace(",", "")) * UNITS[units[i]]
                total += parts[metric]
            ms = float(row[head.index("gpu__time_duration.sum")].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[head.index("gpu__time_duration.sum")]]
            break
    else:
        raise SystemExit("no count_reads_kernel launch in %s" % raw)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data["count_reads_kernel_dram_bytes_per_launch_%s" % config] = int(total)
    data["count_reads_kernel_%s_source" % config] = {"from": note, "kernel": row[name], "ncu_ms": ms, "dram_bytes_read": int(parts["dram__bytes_read.sum"]),
                                                    "dram_bytes_write": int(parts["dram__bytes_write.sum"]), "written_by": "profiles/make_traffic.py"}
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(data, indent=1))


if __name__ == "__main__":
    main()
