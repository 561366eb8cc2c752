"""Turn an ncu summary (tools/ncu_summary.py output) of the dominant kernel into profiles/force_traffic.json, the file
bench.py reads `roofline.traffic` (dram bytes per launch) and the FP64 pipe figure from.
usage: python tools/force_traffic.py <summary.json> <out.json> [source note]"""
import json, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def num(text, table=None):
    val, _, unit = text.partition(" ")
    v = float(val.replace(",", ""))
    return v * table.get(unit, 1.0) if table else v


rows = json.load(open(sys.argv[1]))
r = max(rows, key=lambda d: num(d.get("gpu__time_duration.sum", "0 ms")))
rd, wr = num(r["dram__bytes_read.sum"], UNIT), num(r["dram__bytes_write.sum"], UNIT)
t = r["gpu__time_duration.sum"]
ms = num(t) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(t.split()[-1], 1.0)
out = {"kernel": r["kernel"].replace("void ", "").split("(")[0],
       "source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1],
       "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
       "fp64_pipe_active_pct": num(r.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "0 %")),
       "dram_throughput_pct": num(r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "0 %")),
       "duration_ms_under_ncu": ms}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
