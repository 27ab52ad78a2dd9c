"""DRAM bytes per launch by kernel kind from an ncu --set full report -> profiles/ncu_traffic.json
   python scripts/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/ncu_traffic.json"""
import csv, io, json, subprocess, sys
KIND = [("k_ens_small", "ens_small"), ("k_wide_static", "ens_wide"), ("k_wide_voja", "ens_voja"), ("k_pes", "pes"),
        ("k_lin", "lin"), ("k_decode", "decode"), ("k_cleanup_scan", "cleanup_scan"), ("k_cleanup_pick", "cleanup_pick"),
        ("k_gate", "gate"), ("k_begin", "begin")]
rep, out = sys.argv[1], sys.argv[2]
if rep.endswith(".csv"):      # a committed per-launch summary (scripts/ncu_summary.py) instead of the .ncu-rep itself
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ik, ir, iw, it = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    kind = next((v for k, v in KIND if k in r[ik]), None)
    if kind is None:
        continue
    a = acc.setdefault(kind, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0})
    a["launches"] += 1
    a["dram_bytes"] += float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    a["time_us"] += float(r[it])
res = {k: {"dram_bytes_per_launch": v["dram_bytes"] / v["launches"], "launches_captured": v["launches"],
           "ncu_time_us_per_launch": v["time_us"] / v["launches"]} for k, v in acc.items()}
# whole-timestep DRAM bytes: mean bytes per launch of every kernel x its launches per timestep (a capture window rarely
# starts and ends on a step boundary, so "sum / steps captured" over-counts the ragged ends).  BASELINE configs[1]: two
# levels -> k_ens_small x 2, k_lin x 4 (level 0, level 1, early end-of-step rows, end of step); the deferred-PES fold + clear every 8th step (their
# bytes come from FOLD_MB when the window holds none: profiles/r02c_ncu_pes_summary.csv).
all_acc = {}
for r in rows[2:]:
    name = r[ik].split("(")[0].split("<")[0].replace("void ", "")
    a = all_acc.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
import os
PER_STEP = {"k_ens_small": 2.0, "k_lin": 4.0, "k_pes_fold": 0.125, "k_pes_clear": 0.125, "k_advance": 1.0 / 16}
per_step = {k: v[1] / v[0] * PER_STEP.get(k, 1.0) for k, v in all_acc.items()}
if "k_pes_fold" not in per_step and os.environ.get("FOLD_MB"):
    per_step["k_pes_fold"] = float(os.environ["FOLD_MB"]) * 1e6 * 0.125
res["_step"] = {"dram_bytes_per_timestep": sum(per_step.values()), "trials": int(os.environ.get("B", "1024")),
                "method": "mean DRAM bytes per launch of each kernel x launches per timestep",
                "per_timestep_by_kernel": {k: v for k, v in sorted(per_step.items())},
                "by_kernel": {k: {"launches": v[0], "dram_bytes": v[1]} for k, v in sorted(all_acc.items())}}
res["_source"] = rep.split("/")[-1]
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
