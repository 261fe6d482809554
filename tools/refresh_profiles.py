"""Turns the ncu captures / launch list a GPU run left in gpurun_out/ into the small text files
tracked under profiles/ (summaries, traffic.json, launch shares).
    python tools/refresh_profiles.py [round_tag]        # default r1_final"""
import collections
import csv
import json
import os
import re
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def dram_bytes(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = dict(zip(rows[0], zip(rows[1], rows[2])))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def b(k):
        u, v = d[k]
        return float(v.replace(",", "")) * scale[u]
    return b("dram__bytes_read.sum") + b("dram__bytes_write.sum")


PIPE_KEYS = {
    "issue_active": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex": "l1tex__throughput.avg.pct_of_peak_sustained_active",
}


def pipe_pcts(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = dict(zip(rows[0], rows[2]))
    out = {}
    for k, m in PIPE_KEYS.items():
        if m in d and d[m] not in ("", "n/a"):
            out[k] = round(float(d[m].replace(",", "")), 2)
    return out


traffic, pipes = {}, {}
for f in sorted(os.listdir(G)):
    m = re.match(r"prof_(c2|c3_mini|c3_shard|c3|c4_shard)_(k_[a-z_]+)\.ncu-rep$", f)
    if not m:
        continue
    wl, k = m.groups()
    ncu_summary.main(os.path.join(G, f), os.path.join(P, f"{tag}_{wl}_{k}.txt"))
    if k in ("k_project", "k_project_exact", "k_score", "k_score_mma"):
        traffic.setdefault(wl, {})[k] = dram_bytes(os.path.join(G, f))
        pipes.setdefault(wl, {})[k] = pipe_pcts(os.path.join(G, f))
if traffic:
    traffic["_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none "
                          f"(profiles/{tag}_<workload>_<kernel>.txt)")
    with open(os.path.join(P, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)

if pipes:
    pipes["_source"] = ("ncu --set full --clock-control none, per kernel: % of peak of the issue slots and of each "
                        f"instruction pipe (profiles/{tag}_<workload>_<kernel>.txt)")
    with open(os.path.join(P, "pipes.json"), "w") as f:
        json.dump(pipes, f, indent=1)

lp = os.path.join(G, "launches_bench.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
    h = rows[0]
    ik, iv, ig, iid = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("ID")
    recs = [(int(r[iid]), re.sub(r"^.*?(k_[a-z_]+).*$", r"\1", r[ik]), r[ig], float(r[iv].replace(",", ""))) for r in rows[1:]]
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_* on "
           "`python bench.py --workload c3_shard --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-videos 1`",
           "# per-launch device time (cold-cache, serialised under the profiler: compare SHARES, not absolutes)", ""]
    agg = collections.OrderedDict()
    for _, k, _, v in recs:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    out.append("## all launches of the run (build + warm-up + timed passes + split passes + e2e runs)")
    for k, (n, t) in agg.items():
        out.append(f"{k:14s} n={n:4d} total={t / 1e3:9.1f} us mean={t / n / 1e3:8.2f} us")
    # the device-resident passes: k_unproject, k_project, scoring kernel, k_finalize with the workload's job count
    passes, i = [], 0
    big = max((int(re.sub(r"[^0-9,]", "", g).split(",")[0] or 0) for _, k, g, _ in recs if k == "k_unproject"), default=0)
    while i < len(recs) - 3:
        g0 = int(re.sub(r"[^0-9,]", "", recs[i][2]).split(",")[0] or 0)
        if (recs[i][1] == "k_unproject" and g0 == big and recs[i + 1][1] == "k_project"
                and recs[i + 2][1].startswith("k_score") and recs[i + 3][1] == "k_finalize"):
            passes.append([recs[i + j][3] for j in range(4)])
            score_name = recs[i + 2][1]
            i += 4
        else:
            i += 1
    if passes:
        out += ["", f"## the device-resident passes ({big} jobs): per-pass kernel times and shares"]
        m = [statistics.mean(p[j] for p in passes) / 1e3 for j in range(4)]
        for name, v in zip(("k_unproject", "k_project", score_name, "k_finalize"), m):
            out.append(f"{name:12s} {v:9.2f} us  {100 * v / sum(m):5.1f} % of the pass")
        out.append(f"{'pass total':12s} {sum(m):9.2f} us over {len(passes)} passes")
    open(os.path.join(P, f"{tag}_launches_bench.txt"), "w").write("\n".join(out) + "\n")
for name in ("sanitizer_racecheck.txt", "sanitizer_memcheck.txt"):
    src = os.path.join(G, name)
    if os.path.exists(src):
        with open(src) as f, open(os.path.join(P, f"{tag}_" + name), "w") as o:
            o.writelines(l for l in f if "Host Frame" not in l)
for name, dst in (("bench_default.txt", f"{tag}_bench_default.json"), ("bench_reference.txt", f"{tag}_bench_reference.json"),
                  ("bench_n2.txt", f"{tag}_bench_n2.json"), ("bench_n4.txt", f"{tag}_bench_n4.json"),
                  ("bench_n8.txt", f"{tag}_bench_n8.json"), ("bench_c4_shard.txt", f"{tag}_bench_c4_shard.json"),
                  ("bench_c4_trans.txt", f"{tag}_bench_c4_trans.json"), ("bench_c2.txt", f"{tag}_bench_c2.json")):
    src = os.path.join(G, name)
    if os.path.exists(src):
        lines = [l for l in open(src).read().splitlines() if l.startswith("{")]
        if lines:
            open(os.path.join(P, dst), "w").write(json.dumps(json.loads(lines[-1]), indent=1) + "\n")
print("profiles refreshed:", sorted(os.listdir(P)))
