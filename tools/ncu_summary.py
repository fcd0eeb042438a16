"""Turn an ncu report (.ncu-rep, --set full) into the tracked evidence under profiles/:
  profiles/<tag>_ncu_full_summary.md   per-kernel key metrics
  profiles/kernel_traffic.json         {kernel short name: {"dram_bytes": read+write per launch, ...}} — bench.py reads
                                       it for roofline.traffic
Usage (build container, no GPU needed): python tools/ncu_summary.py gpurun_out/prof_r1c.ncu-rep r01c "<command profiled>"
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def short(name):
    m = re.search(r"(k_[A-Za-z0-9_]+)(<[^(]*>)?", name)
    if not m:
        return name[:60]
    s = m.group(1)
    t = m.group(2) or ""
    for epi in ("EpiScoreT<(bool)1>", "EpiScoreT<(bool)0>", "EpiScoreT<true>", "EpiScoreT<false>", "EpiDz", "EpiStore"):
        if epi in name:
            s += "<" + epi.replace("(bool)1", "save").replace("(bool)0", "nosave").replace("true", "save").replace("false", "nosave") + ">"
            break
    else:
        if t and len(t) < 40:
            s += t
    return s


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full --clock-control none ({tag})", f"# command: {cmd}", f"# report: {rep} (scratch); per-kernel raw metrics below", ""]
    traffic = {}
    for r in body:
        name = r[col["Kernel Name"]]
        lines.append(f"## {name[:140]}")
        vals = {}
        for k in KEYS:
            if k in col:
                vals[k] = (r[col[k]], units[col[k]])
                lines.append(f"{k} = {r[col[k]]} {units[col[k]]}")
        lines.append("")

        def tob(k):
            v, u = vals.get(k, ("0", "byte"))
            return float(v.replace(",", "")) * UNIT.get(u, 1.0)
        d = traffic.setdefault(short(name), {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0})
        d["launches"] += 1
        d["dram_bytes"] += tob("dram__bytes_read.sum") + tob("dram__bytes_write.sum")
        tv, tu = vals.get("gpu__time_duration.sum", ("0", "us"))
        d["time_us"] += float(tv.replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(tu, 1.0)
    for d in traffic.values():
        d["dram_bytes"] /= d["launches"]
        d["time_us"] /= d["launches"]
    for k, d in traffic.items():
        d["source"] = f"profiles/{tag}_ncu_full_summary.md"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.md"), "w").write("\n".join(lines))
    # merge: kernels of other captures (other tags) keep their entries
    tpath = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    try:
        merged = json.load(open(tpath))
    except Exception:
        merged = {}
    merged.update(traffic)
    json.dump(merged, open(tpath, "w"), indent=1)
    for k, d in traffic.items():
        print(k, d)


if __name__ == "__main__":
    main()
