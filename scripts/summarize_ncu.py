"""Turns ncu captures (brought back in gpurun_out/) into the small text summaries committed under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/r2x_launches.csv > profiles/r2_launches_ba_100k_5k.csv
  python scripts/summarize_ncu.py full gpurun_out/r2x_full.ncu-rep [more.ncu-rep ...] > profiles/r2_ncu_full.csv
"""
import csv
import io
import subprocess
import sys

METRICS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
           ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
           ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
           ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
           ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
           ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
           ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_pct"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
           ("smsp__inst_executed.sum", "warp_insts")]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = {}
    for r in rows[hdr + 1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        us = v / 1e3 if r[iu] in ("ns", "nsecond") else (v if r[iu] in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,total_us,share_pct,avg_us")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('"%s",%d,%.1f,%.1f,%.2f' % (k[:150], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))


def full(paths):
    out = []
    for p in paths:
        txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        seen = set()
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            if name in seen:
                continue
            seen.add(name)
            rec = {"kernel": name[:90], "capture": p.split("/")[-1]}
            for m, short in METRICS:
                if m in hdr:
                    rec[short] = "%s %s" % (r[hdr.index(m)], units[hdr.index(m)])
            st = sorted(((float(r[hdr.index(h)] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                         for h in stall), reverse=True)[:5]
            rec["top_stalls_per_issue"] = "; ".join("%s %.2f" % (b, a) for a, b in st)
            out.append(rec)
    keys = ["kernel", "capture"] + [s for _, s in METRICS] + ["top_stalls_per_issue"]
    w = csv.DictWriter(sys.stdout, keys)
    w.writeheader()
    for rec in out:
        w.writerow({k: rec.get(k, "") for k in keys})


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
