"""DRAM bytes per launch by kernel kind from an ncu --set full report -> profiles/ncu_traffic.json
   python scripts/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/ncu_traffic.json"""
import csv, io, json, subprocess, sys
KIND = [("k_ens_small", "ens_small"), ("k_wide_static", "ens_wide"), ("k_wide_voja", "ens_voja"), ("k_pes", "pes"),
        ("k_lin", "lin"), ("k_decode", "decode"), ("k_cleanup_scan", "cleanup_scan"), ("k_cleanup_pick", "cleanup_pick"),
        ("k_gate", "gate"), ("k_begin", "begin")]
rep, out = sys.argv[1], sys.argv[2]
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
# whole-timestep DRAM bytes: every kernel of the capture / number of timesteps captured (k_gate runs once per step; the
# deferred-PES fold runs every 8th step, so captures should span a multiple of 8 steps or carry the fold's share)
all_acc = {}
for r in rows[2:]:
    name = r[ik].split("(")[0].split("<")[0]
    a = all_acc.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
import os
n_steps = int(os.environ.get("CAPTURED_STEPS", "0")) or max(1, all_acc.get("k_gate", [1])[0])
res["_step"] = {"dram_bytes_per_timestep": sum(v[1] for v in all_acc.values()) / n_steps, "timesteps_captured": n_steps,
                "trials": int(os.environ.get("B", "1024")),
                "by_kernel": {k: {"launches": v[0], "dram_bytes": v[1]} for k, v in sorted(all_acc.items())}}
res["_source"] = rep.split("/")[-1]
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
