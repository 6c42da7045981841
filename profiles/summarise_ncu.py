"""Turn ncu CSV exports into the small JSON summaries kept under profiles/ (the .ncu-rep files stay in gpurun_out/).

  python profiles/summarise_ncu.py launches <launches.csv> <out.json> "<command that was profiled>"
  python profiles/summarise_ncu.py full <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> <out.json> [reads per launch]
"""
import collections
import csv
import json
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"\(IndexView.*", "", name)
    return name if len(name) < 90 else name[:87] + "..."


def launches(path, out, command):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    h = {n: i for i, n in enumerate(rows[0])}
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[h["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[h["Metric Value"]].replace(",", ""))
        unit = r[h["Metric Unit"]]
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        k = short(r[h["Kernel Name"]])
        tot[k][0] += 1
        tot[k][1] += ms
    total = sum(v[1] for v in tot.values())
    ks = [{"kernel": k, "launches": v[0], "total_ms": round(v[1], 3), "share": round(v[1] / total, 4)}
          for k, v in sorted(tot.items(), key=lambda x: -x[1][1])]
    json.dump({"command": command, "note": "cold-cache, serialised per-launch times: compare shares, not absolutes",
               "total_ms": round(total, 3), "kernels": ks}, open(out, "w"), indent=1)
    for k in ks[:8]:
        print(k)


KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def full(path, out, reads=None):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    h = {n: i for i, n in enumerate(hdr)}
    res = []
    to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for r in rows[2:]:
        d = {"kernel": short(r[h["Kernel Name"]])}
        for k in KEEP:
            if k in h:
                d[k] = [r[h[k]], units[h[k]]]
        try:
            b = sum(float(d[k][0].replace(",", "")) * to_bytes[d[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            d["dram_bytes"] = b
            if reads:
                d["dram_bytes_per_read"] = b / reads
        except Exception:
            pass
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print(d["kernel"], d.get("gpu__time_duration.sum"), d.get("dram_bytes"), d.get("dram_bytes_per_read"))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3], float(sys.argv[4]) if len(sys.argv) > 4 else None)
