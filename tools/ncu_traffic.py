"""DRAM traffic per launch of the roofline entries bench.py reports, from `ncu --set full` captures.
    python tools/ncu_traffic.py profiles/ncu_traffic.json "ENTRY NAME=path/to/capture.ncu-rep" ...
ENTRY NAME is a `kernel` string of bench.py's `rooflines` (e.g. "conv_fprop 1x1/1 64->256 @56"); the
capture holds ONE launch of the entry's main kernel (ncu -k regex:... -c 1). Writes / updates the JSON
bench.py reads (`traffic` = dram__bytes_read.sum + dram__bytes_write.sum), plus a few pipe metrics.
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6,
         "nsecond": 1e-9, "ms": 1e-3, "msecond": 1e-3, "%": 1, "": 1}


def metrics(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    get = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            get[h] = float(v.replace(",", "")) * UNITS.get(u, 1)
        except ValueError:
            pass
    return get


def main():
    dest = Path(sys.argv[1])
    data = json.loads(dest.read_text()) if dest.exists() else {}
    for spec in sys.argv[2:]:
        name, rep = spec.split("=", 1)
        m = metrics(rep)
        data[name] = dict(
            dram_bytes=m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0),
            dram_read_bytes=m.get("dram__bytes_read.sum"), dram_write_bytes=m.get("dram__bytes_write.sum"),
            duration_us=m.get("gpu__time_duration.sum", 0) * 1e6,
            l2_hit_pct=m.get("lts__t_sector_hit_rate.pct"),
            issue_active_pct=m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            tensor_pipe_pct=m.get("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"),
            capture=str(rep))
        print(name, data[name])
    dest.write_text(json.dumps(data, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
