"""Top stall locations + key metrics of an .ncu-rep (first kernel).  python tools/ncu_top.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
        "smsp__cycles_elapsed.avg", "gpc__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        print(f"{h:70s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
i_src, i_s = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try:
        data.append((int(r[i_s]), r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "sass rows", len(data))
agg = {}
for s_, r in data:
    for i, h in stall_cols:
        agg[h] = agg.get(h, 0) + (int(r[i]) if r[i] else 0)
print("stall mix:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for s_, r in sorted(data, key=lambda d: -d[0])[:n]:
    st = sorted([(int(r[i]) if r[i] else 0, h) for i, h in stall_cols], reverse=True)[:2]
    print(f"{s_:6d} {100*s_/max(tot,1):5.1f}%  {r[i_src][:80]:80s} {st}")
