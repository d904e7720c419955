#!/usr/bin/env python
"""profiles/k_trace_dram_bytes_per_launch.json from an ncu capture (no hand-entered numbers).

    python tools/ncu_traffic.py <raw.csv of `ncu -i x.ncu-rep --page raw --csv`> <stderr of the captured run with TGB_TRACE_BOUNCES=1> <launch index> [out.json]

The captured command is `ncu --set full -k regex:k_trace -s S -c 1 python bench.py ...` run with TGB_TRACE_BOUNCES=1: launch S of
k_trace is wavefront iteration S of the first render call, whose sizes the library prints ("after iter S-1: next n ...").
traffic = dram__bytes_read.sum + dram__bytes_write.sum of that launch; algorithmic bytes = 48 B x the queries it traversed."""
import csv
import json
import re
import sys


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def main():
    raw, log, launch = sys.argv[1], sys.argv[2], int(sys.argv[3])
    out = sys.argv[4] if len(sys.argv) > 4 else "profiles/k_trace_dram_bytes_per_launch.json"
    rows = list(csv.reader(open(raw)))
    d = dict(zip(rows[0], zip(rows[1], rows[2])))
    rd = float(d["dram__bytes_read.sum"][1].replace(",", ""))*unit_scale(d["dram__bytes_read.sum"][0])
    wr = float(d["dram__bytes_write.sum"][1].replace(",", ""))*unit_scale(d["dram__bytes_write.sum"][0])
    dur = float(d["gpu__time_duration.sum"][1].replace(",", ""))
    queries = None
    for ln in open(log, errors="ignore"):
        m = re.match(r"after iter (\d+): next n (\d+) \(survivors (\d+), new (\d+), to traverse (\d+)\)", ln)
        if m and int(m.group(1)) == launch - 1:
            queries = int(m.group(5)) + int(m.group(4))
            break
    if launch == 0:
        queries = None
    js = {"kernel": d.get("Kernel Name", ("", "k_trace"))[1] if "Kernel Name" in d else "k_trace", "launch_index": launch,
          "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "duration_us": dur,
          "queries_in_launch": queries, "algorithmic_bytes": None if queries is None else 48*queries,
          "traffic_over_algorithmic": None if not queries else (rd + wr)/(48.0*queries),
          "source": "%s (ncu --set full, launch %d of k_trace) + %s" % (raw, launch, log)}
    json.dump(js, open(out, "w"), indent=1)
    print(json.dumps(js))


if __name__ == "__main__":
    main()
