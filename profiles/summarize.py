"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (tracked).

  python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/launches_rNN.txt
  python profiles/summarize.py raw gpurun_out/prof.ncu-rep       > profiles/ncu_conv_rNN.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("gd::<unnamed>::", "").replace("void ", "")[:70]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"total {T:.1f} us over {sum(cnt.values())} launches (ncu per-launch times are serialised/cold: compare SHARES)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[k]:4d}  avg={v / cnt[k]:8.1f} us  {k}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name", "")[:100])
        for i, h in enumerate(hdr):
            if any(h == k for k in KEYS):
                print(f"  {h} = {r[i]} {units[i]}")
        print()




def source(path, which=0, top=45):
    """Top SASS instructions by warp-stall samples of the `which`-th kernel in the report (needs -lineinfo capture)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    b = blocks[which]
    h = b["hdr"]
    si = h.index("# Samples")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si] or 0) for r in b["rows"])
    print("kernel:", b["name"][:90], "total samples", tot)
    order = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][si] or 0))[:top]
    for i in sorted(order):
        r = b["rows"][i]
        st = sorted(((int(r[c] or 0), h[c][6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{i:6d} {int(r[si]):7d} {100 * int(r[si]) / max(tot, 1):5.1f}%  {r[1].strip()[:70]:70s} {st}")


if __name__ == "__main__":
    if sys.argv[1] == "source":
        source(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0, int(sys.argv[4]) if len(sys.argv) > 4 else 45)
    else:
        {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
