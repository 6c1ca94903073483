"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/launches_r01_summary.json
  python tools/ncu_summary.py full gpurun_out/prof_glm_r01.ncu-rep profiles/glm_full_r01.json
"""
import csv, io, json, subprocess, sys
from collections import OrderedDict


def launches(path, out):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            if unit in ("us", "usecond"):
                v *= 1e3
            elif unit in ("ms", "msecond"):
                v *= 1e6
            rows.append((r["Kernel Name"].split("(")[0], v))
    agg = OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in rows)
    summ = {"command": "python bench.py --steps 2 --warmup 3 (under ncu --metrics gpu__time_duration.sum --clock-control none)",
            "note": "per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes",
            "launches_captured": len(rows), "total_ms": tot / 1e6,
            "by_kernel": [{"kernel": k, "launches": n, "total_ms": t / 1e6, "share": t / tot, "avg_us": t / n / 1e3}
                          for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    json.dump(summ, open(out, "w"), indent=1)
    print(json.dumps(summ["by_kernel"][:8], indent=1))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "lts__t_bytes.sum", "sm__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"]


def _raw_page(path):
    """the raw page of a report; `path` may already be that page as CSV (reports stay on the GPU box when they are large)"""
    if path.endswith(".csv"):
        return open(path).read()
    return subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout


def full(path, out):
    txt = _raw_page(path)
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    res = []
    for row in data:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        item = {"kernel": d.get("Kernel Name", "").split("(")[0], "id": d.get("ID")}
        for k in hdr:
            if k in KEYS or "fp64" in k or ("stalled" in k and "ratio" in k and "per_issue_active" in k) or \
                    (k.startswith("sm__inst_executed_pipe_") and k.endswith(".avg.pct_of_peak_sustained_active")) or \
                    (k.startswith("sm__pipe_") and k.endswith("cycles_active.avg.pct_of_peak_sustained_active")):
                try:
                    item[k] = float(d[k].replace(",", ""))
                except Exception:
                    item[k] = d[k]
                if u.get(k):
                    item[k + "__unit"] = u[k]
        res.append(item)
    json.dump(res, open(out, "w"), indent=1)
    for it in res:
        print(it["kernel"], {k: v for k, v in it.items() if k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread")})


def traffic_metrics(path, out):
    """the same from a `--metrics ...dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file` capture (one row per metric)"""
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    vals, kern = {}, ""
    for r in csv.DictReader(io.StringIO("".join(lines))):
        vals[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
        kern = r["Kernel Name"].split("(")[0]
        grid, block = r["Grid Size"], r["Block Size"]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd, wr = (vals[k][0] * scale[vals[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    t, tu = vals["gpu__time_duration.sum"]
    res = {"kernel": kern, "grid": grid, "block": block, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "l2_bytes": vals.get("lts__t_bytes.sum", (None, ""))[0],
           "duration_ms_under_ncu": t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, 1e-6),
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none, one launch: " + path}
    json.dump(res, open(out, "w"), indent=1)
    print(res)


def traffic(path, out):
    """dram__bytes_read.sum + dram__bytes_write.sum of the captured kernel (one launch) -> the file bench.py reads for roofline.traffic"""
    txt = _raw_page(path)
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, row = rd[0], rd[1], rd[2]
    d, u = dict(zip(hdr, row)), dict(zip(hdr, units))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(k):
        return float(d[k].replace(",", "")) * scale[u[k]]
    res = {"kernel": d["Kernel Name"].split("(")[0], "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
           "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
           "duration_ms_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u["gpu__time_duration.sum"], 1e-6),
           "source": "ncu --set full --clock-control none, one launch: " + path}
    json.dump(res, open(out, "w"), indent=1)
    print(res)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic, "traffic_metrics": traffic_metrics}[sys.argv[1]](sys.argv[2], sys.argv[3])
