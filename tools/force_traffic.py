"""Turn an ncu --set full capture of the dominant kernel into profiles/force_traffic.json, the calibration file bench.py
reads its static ncu figures from (`roofline.traffic` = DRAM bytes per launch, the FP64 pipe figure, and the L1TEX
data-pipe wavefronts per listed pair behind the primary `roofline`).
usage: python tools/force_traffic.py <capture.ncu-rep> <out.json> <listed pairs in the captured launch> <source note> <commit>"""
import csv, json, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}

rep, out_path, listed = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]


def col(r, name):
    i = hdr.index(name)
    return float(r[i].replace(",", "")), units[i]


r = max(rows[2:], key=lambda q: col(q, "gpu__time_duration.sum")[0] * TIME.get(col(q, "gpu__time_duration.sum")[1], 1.0))
rd, u1 = col(r, "dram__bytes_read.sum")
wr, u2 = col(r, "dram__bytes_write.sum")
t, ut = col(r, "gpu__time_duration.sum")
wf_per_sm = col(r, "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg")[0] if "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg" in hdr else None
n_sm = col(r, "launch__sm_count")[0] if "launch__sm_count" in hdr else 148.0
out = {"kernel": r[hdr.index("Kernel Name")].replace("void ", "").split("(")[0],
       "source": sys.argv[4] if len(sys.argv) > 4 else rep, "commit": sys.argv[5] if len(sys.argv) > 5 else "",
       "dram_bytes_read": rd * UNIT.get(u1, 1.0), "dram_bytes_write": wr * UNIT.get(u2, 1.0),
       "dram_bytes_per_launch": rd * UNIT.get(u1, 1.0) + wr * UNIT.get(u2, 1.0),
       "fp64_pipe_active_pct": col(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")[0],
       "dram_throughput_pct": col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")[0],
       "duration_ms_under_ncu": t * TIME.get(ut, 1.0),
       "l1tex_data_pipe_lsu_wavefronts_pct": col(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")[0],
       "l1tex_lsu_writeback_active_pct": col(r, "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed")[0],
       "l1tex_data_pipe_lsu_wavefronts_mem_lgds_per_sm": wf_per_sm, "listed_pairs_in_capture": listed,
       "l1tex_wavefronts_per_listed_pair": (wf_per_sm * n_sm / listed) if wf_per_sm and listed > 0 else None,
       "l1_sector_hit_rate_pct": col(r, "l1tex__t_sector_hit_rate.pct")[0], "l2_sector_hit_rate_pct": col(r, "lts__t_sector_hit_rate.pct")[0],
       "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")[0]}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out))
