"""`ncu --set full` captures of the kernels behind bench.py's roofline entries, and their summary.

    python tools/ncu_rooflines.py capture OUTDIR [batch]   (GPU box: one ncu run per entry)
    python tools/ncu_rooflines.py parse OUTDIR [JSON]      (anywhere ncu is installed)

`parse` writes profiles/ncu_traffic.json (entry -> DRAM bytes per launch, duration, tensor-pipe
utilisation; bench.py copies `dram_bytes` into `roofline.traffic`) and prints a markdown table
for profiles/. Numbers taken under the profiler are never bench values: only the byte counts and
pipe utilisation are used.
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
# entry name in bench.kernel_rooflines -> (file stem, kernel regex, kernels per call)
ENTRIES = [
    ("conv_fprop 1x1 64->256 @56", "fprop_1x1_64_256_56", "conv_tma_fwd_kernel", 1),
    ("conv_dgrad 1x1 64->256 @56", "dgrad_1x1_64_256_56", "conv_tma_fwd_kernel", 1),
    ("conv_wgrad 1x1 64->256 @56", "wgrad_1x1_64_256_56", "conv_tma_wgrad_kernel", 1),
    ("conv_fprop 1x1 256->1024 @14", "fprop_1x1_256_1024_14", "conv_tma_fwd_kernel", 1),
    ("conv_fprop 3x3 256->256 @14", "fprop_3x3_256_256_14", "conv_tma_fwd_kernel", 1),
    ("conv_wgrad 3x3 256->256 @14", "wgrad_3x3_256_256_14", "conv_tma_wgrad_kernel", 1),
    ("conv_fprop 3x3 512->512 @7", "fprop_3x3_512_512_7", "conv_tma_fwd_kernel", 1),
    ("bn_apply+relu (statistics from the conv epilogue)", "bn_apply", "bn_apply_fast_kernel", 1),
    ("bn_forward_train(stats+apply+relu)", "bn_forward", "bn_stats_kernel|bn_apply_fast_kernel", 2),
    ("bn_backward(reduce+apply, relu fused)", "bn_backward", "bn_bwd_reduce_kernel|bn_bwd_apply", 2),
    ("eltwise_add_relu", "eltwise", "eltwise_fwd_kernel", 1),
    ("maxpool_forward k3s2", "maxpool_fwd", "maxpool_fwd_k3s2", 1),
    ("maxpool_backward k3s2", "maxpool_bwd", "maxpool_bwd_k3s2", 1),
]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6,
        "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6, "%": 1, "": 1}


def capture(out: Path, batch: int):
    out.mkdir(parents=True, exist_ok=True)
    for name, stem, regex, per_call in ENTRIES:
        cmd = ["ncu", "--set", "full", "--clock-control", "none", "-k",
               f"regex:{regex}", "--launch-skip", str(per_call), "-c", str(per_call), "-f", "-o",
               str(out / stem), sys.executable, str(ROOT / "tools" / "rooflines.py"), str(batch), name]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        (out / f"{stem}.log").write_text(r.stdout[-2000:] + r.stderr[-2000:])
        print(f"{stem}: rc={r.returncode}", flush=True)


def parse(out: Path):
    table, md = {}, ["| roofline entry | kernel(s) | time under ncu (us) | DRAM read (MB) | DRAM write (MB) | "
                     "DRAM busy % | tensor pipe % | L2 hit % | issue active % |", "|---|---|---|---|---|---|---|---|---|"]
    for name, stem, _regex, _n in ENTRIES:
        rep = out / f"{stem}.ncu-rep"
        if not rep.exists():
            continue
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        tot = dict(us=0.0, rd=0.0, wr=0.0)
        kernels, pct = [], {}
        for r in rows[2:]:
            g = lambda k: (float(r[hdr.index(k)].replace(",", "") or 0) * UNIT.get(units[hdr.index(k)], 1)
                           if k in hdr else 0.0)
            kernels.append(r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", ""))
            tot["us"] += g(METRICS[0]); tot["rd"] += g(METRICS[1]); tot["wr"] += g(METRICS[2])
            for k in METRICS[3:]:
                pct.setdefault(k, []).append(g(k))
        table[name] = dict(kernels=kernels, dram_bytes=tot["rd"] + tot["wr"], dram_read_bytes=tot["rd"],
                           dram_write_bytes=tot["wr"], duration_us_under_ncu=tot["us"],
                           tensor_pipe_pct=max(pct[METRICS[3]]), dram_pct_of_peak=max(pct[METRICS[4]]),
                           l2_hit_pct=sum(pct[METRICS[5]]) / len(kernels),
                           issue_active_pct=max(pct[METRICS[6]]), source=f"{out.name}/{stem}.ncu-rep")
        md.append(f"| {name} | {' + '.join(kernels)} | {tot['us']:.1f} | {tot['rd'] / 1e6:.1f} | "
                  f"{tot['wr'] / 1e6:.1f} | {max(pct[METRICS[4]]):.1f} | {max(pct[METRICS[3]]):.1f} | "
                  f"{table[name]['l2_hit_pct']:.1f} | {max(pct[METRICS[6]]):.1f} |")
    dst = Path(sys.argv[3]) if len(sys.argv) > 3 else ROOT / "profiles" / "ncu_traffic.json"
    dst.write_text(json.dumps(table, indent=1))
    print("\n".join(md))


if __name__ == "__main__":
    mode, out = sys.argv[1], Path(sys.argv[2])
    if mode == "capture":
        capture(out, int(sys.argv[3]) if len(sys.argv) > 3 else 256)
    else:
        parse(out)
