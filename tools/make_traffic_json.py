"""profiles/score_bounds_traffic.json from an `ncu --set full` capture of score_bounds at the bench's
launch size (bench.py reads it into roofline.traffic).  Records the commit the capture was taken from.
    python tools/make_traffic_json.py gpurun_out/r2b_score_bounds.ncu-rep "<capture command>"
"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
      "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def num(k):
    v, u = d[k]
    return float(v.replace(",", "")) * SC.get(u, 1.0)


head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
dirty = subprocess.run(["git", "status", "--porcelain", "deep-sfm-revisited_b200/csrc"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
out = {"kernel": d["Kernel Name"][0].split("(")[0].replace("void ", "tv5::"),
       "launch": "bench.py workload: 256 pairs x 10,000 correspondences, ~2.83e6 hypotheses (one launch per step)",
       "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
       "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
       "duration_s_under_ncu": num("gpu__time_duration.sum"),
       "fma_pipe_pct": float(d["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][0]),
       "grid": d["launch__grid_size"][0], "registers": d["launch__registers_per_thread"][0],
       "commit": head + (" + uncommitted changes in csrc/" if dirty else ""),
       "report": os.path.basename(rep), "source": cmd}
json.dump(out, open(os.path.join(ROOT, "profiles", "score_bounds_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
