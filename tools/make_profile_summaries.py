#!/usr/bin/env python
"""Turn the ncu outputs under gpurun_out/ into the text summaries committed under profiles/ (run where ncu can read reports)."""
import collections, csv, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

METRICS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"] + \
          ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k for k in
           ("long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "not_selected", "branch_resolving", "no_instruction",
            "dispatch_stall", "mio_throttle", "lg_throttle", "barrier")]


def launches(csvfile, title, out):
    rows = list(csv.reader(l for l in open(csvfile) if not l.startswith("==")))
    hdr = rows[0]; ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except Exception:
            continue
        name = r[ik].split("(")[0]
        name = name.replace("void ", "").replace("<unnamed>::", "")[-70:]
        tot[name] += v; cnt[name] += 1
    s = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"# total captured device time {s / 1e6:.2f} ms over {sum(cnt.values())} launches\n\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total_us':>12s} {'share':>7s}\n")
        for k, v in tot.most_common(40):
            f.write(f"{k:72s} {cnt[k]:8d} {v / 1e3:12.1f} {100 * v / s:6.1f}%\n")


def _raw_rows(rep):
    """raw-metric table of a capture: from the .raw.csv exported on the GPU box (tools/profile.sh) or from the report itself"""
    if rep.endswith(".csv"):
        return list(csv.reader(open(rep)))
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(txt.splitlines()))


def traffic_json(rep, out, nparticles):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the pair kernels -> profiles/ncu_traffic.json (read by bench.py)"""
    import json
    rows = _raw_rows(rep)
    hdr, units = rows[0], rows[1]
    def tobytes(v, u):
        v = float(v)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    kern = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        b = tobytes(r[ir], units[ir]) + tobytes(r[iw], units[iw])
        key = "k_density" if "k_density" in name else ("k_force" if "k_force" in name else None)
        if key and key not in kern:
            kern[key] = int(b)
    json.dump({"source": "ncu --set full --clock-control none, bench.py --steps 1 --warmup 3 (tools/profile.sh); dram__bytes_read.sum + dram__bytes_write.sum per launch",
               "workload_particles": nparticles, "kernels": kern}, open(out, "w"), indent=1)


def raw(rep, title, out, mintime_ms=0.2):
    rows = _raw_rows(rep)
    hdr, units = rows[0], rows[1]
    seen = set()
    with open(out, "w") as f:
        f.write(f"# {title}\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            t = float(r[hdr.index("gpu__time_duration.sum")])
            key = name.split("(")[0]
            if t < mintime_ms or key in seen:
                continue
            seen.add(key)
            f.write(f"== {name[:110]}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"  {m.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', ''):78s} {r[i]:>18s} {units[i]}\n")
            f.write("\n")


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    P = lambda n: os.path.join(ROOT, "profiles", f"{TAG}_{n}")
    if os.path.exists(os.path.join(OUT, "launches.csv")):
        launches(os.path.join(OUT, "launches.csv"), "ncu launch list: bench.py --steps 2 --warmup 3 (turb 128^3, 1 B200); tools/profile.sh", P("launches_turb128.txt"))
    if os.path.exists(os.path.join(OUT, "launches_grav.csv")):
        launches(os.path.join(OUT, "launches_grav.csv"), "ncu launch list: tools/bench_gravity.py 1e6 2 (self-gravitating sphere, 1M particles, 2 derivs calls)", P("launches_gravity1M.txt"))
    pair = os.path.join(OUT, "prof_pair.raw.csv") if os.path.exists(os.path.join(OUT, "prof_pair.raw.csv")) else os.path.join(OUT, "prof_pair.ncu-rep")
    grav = os.path.join(OUT, "prof_grav.raw.csv") if os.path.exists(os.path.join(OUT, "prof_grav.raw.csv")) else os.path.join(OUT, "prof_grav.ncu-rep")
    if os.path.exists(pair):
        raw(pair, "ncu --set full --clock-control none: density / force pair kernels, turb 128^3 (2,097,152 particles, 57 neighbours)", P("pair_kernels_ncu.txt"))
        traffic_json(pair, os.path.join(ROOT, "profiles", "ncu_traffic.json"), 128 ** 3)
    if os.path.exists(grav):
        raw(grav, "ncu --set full --clock-control none: self-gravity kernels, random sphere 1M particles, tree_accuracy 0.5", P("gravity_kernels_ncu.txt"))
