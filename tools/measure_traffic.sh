#!/bin/bash
# roofline.traffic of bench.py's MPPI line, measured in the run's own conditions (VERDICT r01 "weak" 6): DRAM bytes of 16
# consecutive rollout launches through the state ring as ONE ncu range, divided by 16; writes profiles/roofline_traffic.json.
#   gpurun -- bash tools/measure_traffic.sh        (one GPU; needs the built library)
set -e
mkdir -p gpurun_out
ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --csv --log-file gpurun_out/r02_traffic_range.csv python tools/traffic_probe.py > gpurun_out/r02_traffic_probe.log 2>&1
python - <<'PY'
import csv, json, os
rows = [r for r in csv.reader(open("gpurun_out/r02_traffic_range.csv")) if len(r) > 5]
hdr = rows[0]
name, val, unit = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
for r in rows[1:]:
    v = float(r[val].replace(",", ""))
    u = r[unit]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    tot[r[name]] = tot.get(r[name], 0.0) + v * scale
n = 16
path = os.path.join("profiles", "roofline_traffic.json")
old = json.load(open(path)) if os.path.exists(path) else {}
old["mppi_rollout_kernel_dram_bytes_per_launch"] = (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / n
old.setdefault("notes", {})["mppi_rollout_kernel"] = (
    "round 2: ncu --replay-mode range over %d consecutive launches of the rollout phase writing through the 16-buffer state ring "
    "(tools/measure_traffic.sh): %.1f MB read + %.1f MB written in the range, per launch %.2f MB against 12.58 MB algorithmic "
    "(+ 8.4 MB of variates read from the buffer the noise kernel filled, L2-resident between launches)"
    % (n, tot["dram__bytes_read.sum"] / 1e6, tot["dram__bytes_write.sum"] / 1e6, old["mppi_rollout_kernel_dram_bytes_per_launch"] / 1e6))
old["source"] = "mppi: tools/measure_traffic.sh (range replay, final r02 build); rbpf: profiles/r01c_*_ncu_full_selected.csv"
json.dump(old, open("gpurun_out/roofline_traffic.json", "w"), indent=1)
print(json.dumps(old, indent=1))
PY
