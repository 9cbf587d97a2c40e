"""Summarise ncu artefacts into markdown for profiles/ (run in the build container; ncu only reads files).

  python tools/ncu_summary.py launches gpurun_out/launches_r01k.csv [profiles/r01k_traffic.json] > profiles/r01k_launches.md
  python tools/ncu_summary.py report   gpurun_out/prof_tapgemm_r01k.ncu-rep ... > profiles/r01k_ncu_summary.md
"""
from __future__ import annotations

import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
]


def _ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def _num(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return 0.0


def launches(path, json_out=None):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: {"ids": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
    for r in rows[start + 1:]:
        if len(r) != len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("vg::", "")
        a = agg[name]
        a["ids"].add(r[ix["ID"]])
        metric, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], _num(r[ix["Metric Value"]])
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
        if metric.startswith("gpu__time_duration"):
            a["ns"] += val * scale
        elif metric.startswith("dram__bytes_read"):
            a["rd"] += val * scale
        elif metric.startswith("dram__bytes_write"):
            a["wr"] += val * scale
    tot = sum(v["ns"] for v in agg.values())
    print(f"# kernel launch list ({path.split('/')[-1]}): `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`")
    print("\nPer-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live breakdown, not absolutes.\n")
    print("| kernel | launches | total us | share | DRAM read MB / launch | DRAM write MB / launch |\n|---|---:|---:|---:|---:|---:|")
    out = {}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        n = len(v["ids"])
        out[k] = {"launches": n, "us": v["ns"] / 1e3, "dram_read_bytes_per_launch": v["rd"] / n, "dram_write_bytes_per_launch": v["wr"] / n}
        if v["ns"] / tot < 0.0005:
            continue
        print(f"| `{k}` | {n} | {v['ns'] / 1e3:.1f} | {100 * v['ns'] / tot:.1f}% | {v['rd'] / n / 1e6:.2f} | {v['wr'] / n / 1e6:.2f} |")
    print(f"\ntotal {tot / 1e6:.2f} ms over {sum(len(v['ids']) for v in agg.values())} launches (one CFG-batched UNet forward + the step update)")
    if json_out:
        import json
        json.dump(out, open(json_out, "w"), indent=1)


def report(path):
    raw = list(csv.reader(io.StringIO(_ncu(["-i", path, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"\n## {path.split('/')[-1]}  (`ncu --set full --clock-control none --import-source on`)\n")
    for li, r in enumerate(raw[2:]):
        print(f"### launch {li}: `{r[ix['Kernel Name']][:90]}`\n")
        print("| metric | value |\n|---|---:|")
        for key, label in RAW_KEYS:
            if key in ix:
                print(f"| {label} | {r[ix[key]]} {units[ix[key]]} |")
        src = list(csv.reader(io.StringIO(_ncu(["-i", path, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"]))))
        if len(src) < 3:
            continue
        sh = src[1]
        six = {h: i for i, h in enumerate(sh)}
        data = [x for x in src[2:] if len(x) == len(sh)]
        stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(_num(x[six["# Samples"]]) for x in data) or 1.0
        agg = sorted(((sum(_num(x[six[s]]) for x in data), s) for s in stalls), reverse=True)[:6]
        print("\nwarp-state samples: " + ", ".join(f"{s[6:]} {100 * v / tot:.1f}%" for v, s in agg))
        print("\nhottest SASS lines (share of samples, top stall):\n")
        print("| % | SASS | stall |\n|---:|---|---|")
        for x in sorted(data, key=lambda x: -_num(x[six["# Samples"]]))[:10]:
            st = max(((_num(x[six[s]]), s) for s in stalls))
            print(f"| {100 * _num(x[six['# Samples']]) / tot:.1f} | `{x[six['Source']][:70]}` | {st[1][6:]} |")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        print("# ncu summaries\n\nRead with `ncu -i <rep> --page raw|source --csv`; the .ncu-rep files stay in gpurun_out/ (scratch).")
        for p in sys.argv[2:]:
            report(p)
