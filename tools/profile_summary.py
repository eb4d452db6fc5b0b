"""Turns the ncu artefacts of a gpurun call into the tracked summaries under
profiles/:  <tag>_launches.csv (verbatim launch list), <tag>_launch_shares.txt,
<tag>_<kernel>_details.csv (ncu details page of the first captured launch),
<tag>_<kernel>_hot_sass.txt (stall samples per SASS line) and traffic.json
(DRAM bytes per launch, read by bench.py).

usage: python tools/profile_summary.py <tag> <launches.csv> <prof.ncu-rep> <kernel-short-name> <traffic-key>
(traffic-key: the bench.py workload key under which the DRAM bytes per SpMV go
into profiles/traffic.json, e.g. lap27_distinct_f64; a step that runs two
kernels -- R-MAT: row kernel + hub kernel -- is the sum of both)
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")


def ncu_page(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"],
                          capture_output=True, text=True).stdout


def launches(tag, path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = (hdr.index("Kernel Name"), hdr.index("Metric Value"),
                  hdr.index("Metric Unit"))
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, tag + "_launches.csv"), "w") as f:
        f.write(open(path).read())
    with open(os.path.join(OUT, tag + "_launch_shares.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none "
                "(cold-cache, serialised: compare SHARES)\n")
        f.write("# %d launches, %.1f us total\n" % (len(data), tot))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%5d launches %11.1f us %6.2f%% avg %9.1f us  %s\n" % (
                n, t, 100 * t / tot, t / n, k[:110]))


def details(tag, rep, kname, traffic_key):
    rows = list(csv.reader(io.StringIO(ncu_page(rep, "details"))))
    h = rows[0]
    keep = [r for r in rows[1:] if r[h.index("ID")] == "0"]
    with open(os.path.join(OUT, "%s_%s_details.csv" % (tag, kname)), "w") as f:
        w = csv.writer(f)
        w.writerow(["Section", "Metric", "Unit", "Value"])
        for r in keep:
            w.writerow([r[h.index("Section Name")], r[h.index("Metric Name")],
                        r[h.index("Metric Unit")], r[h.index("Metric Value")]])
    raw = list(csv.reader(io.StringIO(ncu_page(rep, "raw"))))
    hdr = raw[0]

    def metric(name, row):
        return float(row[hdr.index(name)].replace(",", ""))
    per = collections.OrderedDict()  # kernel name -> launches
    for row in raw[2:]:
        units = raw[1]
        def in_bytes(name):
            v = metric(name, row)
            u = units[hdr.index(name)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        per.setdefault(row[hdr.index("Kernel Name")][:80], []).append({
            "dram_bytes_read": in_bytes("dram__bytes_read.sum"),
            "dram_bytes_write": in_bytes("dram__bytes_write.sum"),
            "duration_us": metric("gpu__time_duration.sum", row) *
            {"us": 1, "ms": 1e3, "ns": 1e-3}[units[hdr.index("gpu__time_duration.sum")]],
        })
    t = {
        "kernel": " + ".join(per.keys()),
        "launches_captured": sum(len(v) for v in per.values()),
        "dram_bytes_per_launch": sum(
            sum(p["dram_bytes_read"] + p["dram_bytes_write"] for p in v) / len(v)
            for v in per.values()),
        "kernel_us": {k: sum(p["duration_us"] for p in v) / len(v)
                      for k, v in per.items()},
        "per_launch": [dict(p, kernel=k) for k, v in per.items() for p in v],
        "source": "ncu --set full --clock-control none, %s" % os.path.basename(rep),
        "round": tag.split("_")[0],
    }
    tpath = os.path.join(OUT, "traffic.json")
    try:
        allt = json.load(open(tpath))
    except Exception:
        allt = {}
    allt[traffic_key] = t
    json.dump(allt, open(tpath, "w"), indent=1)
    # gather / atomic counters of the first captured launch (north star: L2 hit
    # rate of the x gathers, atomic / replay counters)
    wanted = [
        "lts__t_sector_hit_rate.pct",
        "lts__t_sector_op_read_hit_rate.pct",
        "lts__t_sector_op_red_hit_rate.pct",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum",
        "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__time_duration.sum",
    ]
    first = raw[2]
    counters = {}
    for name in wanted:
        if name in hdr:
            counters[name] = {"value": first[hdr.index(name)],
                              "unit": raw[1][hdr.index(name)]}
    json.dump(counters, open(os.path.join(
        OUT, "%s_%s_counters.json" % (tag, kname)), "w"), indent=1)
    # hot SASS
    src = list(csv.reader(io.StringIO(ncu_page(rep, "source"))))
    for idx, r in enumerate(src[:10]):
        if "Source" in r:
            break
    hdr = src[idx]
    S, W, I = (hdr.index("Source"),
               hdr.index("Warp Stall Sampling (All Samples)"),
               hdr.index("Instructions Executed"))
    stall = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not" not in x]
    body = src[idx + 1:]
    agg = {}
    lines = []
    for n, r in enumerate(body):
        try:
            lines.append((n, float(r[W]), float(r[I]), r[S]))
        except Exception:
            continue
        for i in stall:
            try:
                agg[hdr[i]] = agg.get(hdr[i], 0) + float(r[i])
            except Exception:
                pass
    tot = sum(l[1] for l in lines) or 1
    with open(os.path.join(OUT, "%s_%s_hot_sass.txt" % (tag, kname)), "w") as f:
        f.write("# stall reasons (samples): %s\n" % sorted(
            agg.items(), key=lambda kv: -kv[1])[:8])
        f.write("# top SASS lines by warp-stall samples\n")
        for n, w_, i_, s_ in sorted(lines, key=lambda l: -l[1])[:25]:
            f.write("%5d %6.2f%% samples %10.0f executions  %s\n" % (
                n, 100 * w_ / tot, i_, s_[:100]))


if __name__ == "__main__":
    tag, lpath, rep, kname, traffic_key = sys.argv[1:6]
    os.makedirs(OUT, exist_ok=True)
    launches(tag, lpath)
    details(tag, rep, kname, traffic_key)
    print(open(os.path.join(OUT, tag + "_launch_shares.txt")).read()[:1500])
    t = json.load(open(os.path.join(OUT, "traffic.json")))[traffic_key]
    print(traffic_key, t["dram_bytes_per_launch"], t["kernel_us"])
