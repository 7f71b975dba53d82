"""ncu report of profiles/ncu_targets.py -> profiles/ncu_summary_r02.json (read by bench.py) and a readable table.

  python profiles/ncu_summarize.py gpurun_out/ncu_targets_r02.ncu-rep [batch]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-9, "usecond": 1e-6,
         "msecond": 1e-3, "second": 1.0, "%": 1.0, "": 1.0}


def val(r, name):
    if name not in ix or r[ix[name]] in ("", "n/a"):
        return None
    return float(r[ix[name]].replace(",", "")) * SCALE.get(units[ix[name]], 1.0)


# kernels per target, in launch order (profiles/ncu_targets.py)
ORDER = [("conv_gn_256_256_b64", ["conv_igemm_kernel"]), ("conv_gn_128_128_b64", ["conv_igemm_kernel"]),
         ("conv_plain_128_128_b64", ["conv_igemm_kernel"]),
         ("gn_bwd_128_b64", ["gn_bwd_stats_kernel", "gn_finalize_kernel", "gn_bwd_apply_kernel"]),
         ("gn_apply_pool_256_b64", ["gn_apply_kernel"]), ("posterior_b64", ["posterior_kernel"]),
         ("attn_fwd_tc_b64", ["attn_fwd_tc_kernel"]), ("conv_gn_512_256_b64", ["conv_igemm_kernel"]),
         ("conv_gn_256_256_skip512_b64", ["conv_igemm_kernel"]), ("conv_gn_256_256_res_b64", ["conv_igemm_kernel"])]
launches = [r for r in data if len(r) == len(hdr)]
pos, summary, lines = 0, {}, []
for key, kernels in ORDER:
    ent = {"batch": batch, "kernels": [], "duration_us": 0.0, "dram_bytes_per_launch": 0.0,
           "source": f"ncu --set full --clock-control none, {os.path.basename(rep)} (profiles/ncu_targets.py, batch {batch}), "
                     "dram__bytes_read.sum + dram__bytes_write.sum"}
    for k in kernels:
        while pos < len(launches) and k not in launches[pos][ix["Kernel Name"]]:
            pos += 1
        if pos >= len(launches):
            break
        r = launches[pos]
        pos += 1
        dur = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        ent["kernels"].append(r[ix["Kernel Name"]][:90])
        ent["duration_us"] += (dur or 0.0) * 1e6
        ent["dram_bytes_per_launch"] += (rd or 0.0) + (wr or 0.0)
        tp = val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        if tp is not None and "conv" in k or "attn" in k:
            ent["tensor_pipe_active_pct"] = tp
        for nm, col in (("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                        ("registers", "launch__registers_per_thread"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct")):
            v = val(r, col)
            if v is not None:
                ent[nm] = v
    summary[key] = ent
    lines.append(f"{key:30s} {ent['duration_us']:9.1f} us  dram {ent['dram_bytes_per_launch'] / 1e6:9.1f} MB  "
                 f"({ent['dram_bytes_per_launch'] / max(ent['duration_us'], 1e-9) / 1e3:7.1f} GB/s)  tensor pipe "
                 f"{ent.get('tensor_pipe_active_pct', float('nan')):5.1f} %  dram {ent.get('dram_throughput_pct', float('nan')):5.1f} %  "
                 f"regs {ent.get('registers', 0):.0f}")
with open(os.path.join(ROOT, "profiles", "ncu_summary_r02.json"), "w") as f:
    json.dump(summary, f, indent=1)
print("\n".join(lines))
