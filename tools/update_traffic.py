#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report of `tools/stage_times.py --frames F --reps 1 --only ...`:
DRAM bytes (read + write) per frame of the LAST launch of each kernel.  usage: update_traffic.py report.ncu-rep F [out.json]"""
import csv
import io
import json
import os
import subprocess
import sys

rep, frames = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
H, U = rows[0], rows[1]
iname, ird, iwr = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum")


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


last = {}
for r in rows[2:]:
    last[r[iname]] = to_bytes(r[ird], U[ird]) + to_bytes(r[iwr], U[iwr])


def per_frame(sub):
    hit = [v for k, v in last.items() if sub in k]
    return round(hit[0] / frames) if hit else None


t = {"k_frame_pass_shared_dram_bytes_per_frame": per_frame("k_frame_pass_mf<3>"),
     "k_frame_pass_plain_dram_bytes_per_frame": per_frame("k_frame_pass_mf<0>"),
     "k_frame_pass_mask_only_dram_bytes_per_frame": per_frame("k_frame_pass_mf<4>"),
     "k_frame_pass_whole_dram_bytes_per_frame": per_frame("k_frame_pass_mf<1>"),
     "k_blend_px_dram_bytes_per_frame": per_frame("k_blend_px"),
     "k_mask_b_dram_bytes_per_frame": per_frame("k_mask_b"),
     "k_raster_dram_bytes_per_frame": per_frame("k_raster("),
     "k_search_dram_bytes_per_frame": per_frame("k_search"),
     "source": "ncu --set full --clock-control none, %d-frame launches of tools/stage_times.py --reps 1; dram__bytes_read.sum + "
               "dram__bytes_write.sum of the last launch of each kernel / %d (tools/update_traffic.py)" % (frames, frames)}
if t["k_frame_pass_mask_only_dram_bytes_per_frame"] and t["k_mask_b_dram_bytes_per_frame"]:
    t["ld_mask_dram_bytes_per_frame"] = t["k_frame_pass_mask_only_dram_bytes_per_frame"] + t["k_mask_b_dram_bytes_per_frame"]
try:
    old = json.load(open(out))
except Exception:
    old = {}
old.update({k: v for k, v in t.items() if v is not None})
json.dump(old, open(out, "w"), indent=1)
print(json.dumps(t))
