"""Turn ncu output into the small JSON files kept under profiles/.

  python tools/ncu_summarize.py full   gpurun_out/prof_x.ncu-rep  profiles/ncu_full_rN.json [profiles/ncu_traffic_rN.json]
      one `ncu --set full` capture -> per kernel (averaged over its captured launches): duration, DRAM bytes
      read / written, DRAM / L2 / SM throughput %, pipe utilisation, issue rate, registers, achieved occupancy.
      The optional second file is what bench.py quotes as roofline.traffic: {kernel: DRAM bytes per launch}.
  python tools/ncu_summarize.py launches gpurun_out/launches.csv   profiles/launches_rN_summary.json
      launch list of `ncu --metrics gpu__time_duration.sum` -> per kernel count / total / mean (us) and shares.
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.split("(")[0]
    name = name.replace("<unnamed>::", "")
    return name


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


WANT = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fp64_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "pipe_tensor_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("launch__registers_per_thread", "registers"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
])
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0,
              "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}


def full(rep, out, traffic_out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {}
    for i, h in enumerate(hdr):
        if "Triage" in h:
            continue
        for want in WANT:
            if h == want or h.endswith("." + want):
                col.setdefault(want, i)
    per = defaultdict(lambda: defaultdict(list))
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if len(r) <= ki:
            continue
        k = short(r[ki])
        for want, i in col.items():
            v = num(r[i])
            if v is None:
                continue
            v *= UNIT_SCALE.get(units[i], 1.0) if WANT[want] in ("duration", "dram_read", "dram_write") else 1.0
            per[k][WANT[want]].append(v)
    res = OrderedDict()
    for k, d in per.items():
        e = {"launches_captured": len(d.get("duration", []))}
        for name, vals in d.items():
            e[name + ("_us" if name == "duration" else "_bytes" if name.startswith("dram_r") or name.startswith("dram_w") else "")] = sum(vals) / len(vals)
        if "dram_read_bytes" in e and "dram_write_bytes" in e:
            e["dram_bytes"] = e["dram_read_bytes"] + e["dram_write_bytes"]
            if e.get("duration_us"):
                e["dram_gbs"] = e["dram_bytes"] / (e["duration_us"] * 1e-6) / 1e9
        res[k] = e
    res["_source"] = "ncu --set full --clock-control none, %s (per-launch averages; cold-cache, serialised replays)" % rep
    json.dump(res, open(out, "w"), indent=1)
    if traffic_out:
        t = {k.split("<")[0]: v["dram_bytes"] for k, v in res.items() if isinstance(v, dict) and "dram_bytes" in v}
        t["_source"] = res["_source"]
        json.dump(t, open(traffic_out, "w"), indent=1)
    for k, v in res.items():
        if isinstance(v, dict):
            print("%-34s %9.1f us  dram %7.3f GB  %5.1f%% dram  xu %s fma %s fp64 %s tensor %s" % (
                k[:34], v.get("duration_us", 0), v.get("dram_bytes", 0) / 1e9, v.get("dram_pct", 0),
                v.get("pipe_xu_pct"), v.get("pipe_fma_pct"), v.get("pipe_fp64_pct"), v.get("pipe_tensor_pct")))


def launches(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[h + 1:]:
        if len(r) <= vi or num(r[vi]) is None:
            continue
        us = num(r[vi]) * UNIT_SCALE.get(r[ui], 1e-3)
        k = short(r[ki])
        a = agg.setdefault(k, {"count": 0, "total_us": 0.0})
        a["count"] += 1
        a["total_us"] += us
        total += us
    for a in agg.values():
        a["mean_us"] = a["total_us"] / a["count"]
        a["share"] = a["total_us"] / total if total else 0.0
    res = {"total_us": total, "kernels": OrderedDict(sorted(agg.items(), key=lambda kv: -kv[1]["total_us"])),
           "_source": "ncu --metrics gpu__time_duration.sum --clock-control none, %s" % path}
    json.dump(res, open(out, "w"), indent=1)
    for k, a in res["kernels"].items():
        print("%-40s n=%4d mean %10.1f us total %10.1f us  %5.1f%%" % (k[:40], a["count"], a["mean_us"], a["total_us"], 100 * a["share"]))


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(*sys.argv[2:5])
    else:
        launches(sys.argv[2], sys.argv[3])
