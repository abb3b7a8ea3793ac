"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
  python tools/summarize_ncu.py report   gpurun_out/prof.ncu-rep  > profiles/rNN_top_kernels.md
  python tools/summarize_ncu.py traffic  gpurun_out/prof.ncu-rep  > profiles/rNN_traffic.json   (read by bench.py)
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    acc = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0].replace("void ", "")
        v = float(r[iv].replace(",", ""))
        u = r[hdr.index("Metric Unit")]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)     # -> us
        a = acc.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in acc.values())
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f%% |" % (k, a[0], a[1], 100 * a[1] / tot))
    print("\ntotal %.1f us over %d launches (ncu serialised, cold-cache: compare shares, not absolutes)" % (tot, sum(a[0] for a in acc.values())))
    # bench.py also measures three stages outside the timed chain (deterministic tracks_current, packet builder, light
    # triggers): their kernels are in the same process, so the share of the chain's dominant kernel is given separately
    side = ("k_tracks_current", "k_pkt_", "k_scan3", "k_lt_", "k_table_lookup")
    chain = {k: a for k, a in acc.items() if not k.startswith(side)}
    ctot = sum(a[1] for a in chain.values())
    print("\nkernels of the timed chain only (%.1f us): " % ctot + ", ".join(
        "%s %.1f%%" % (k.split("<")[0], 100 * a[1] / ctot) for k, a in sorted(chain.items(), key=lambda kv: -kv[1][1])[:6]))


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("## %s" % r[ik][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("- %s = %s %s" % (k, r[i], units[i]))
        print()


def traffic(path):
    """Per kernel (first captured launch of each): DRAM bytes and the counters bench.py quotes in its roofline block."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]

    def val(r, k, scale=True):
        if k not in hdr:
            return None
        i = hdr.index(k)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        if scale and "byte" in u:
            v *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
        return v
    res = {"_source": "ncu --set full --clock-control none capture of `python bench.py --steps 2 --warmup 1` (%s), per launch: "
                      "dram__bytes_read.sum + dram__bytes_write.sum" % path.split("/")[-1]}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]
        if name in res:
            continue
        t = val(r, "gpu__time_duration.sum")
        tu = units[hdr.index("gpu__time_duration.sum")].lower()
        ms = t * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(tu, 1.0)
        res[name] = {"dram_bytes": int(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")), "ms": round(ms, 4),
                     "l1tex_throughput_pct": val(r, "l1tex__throughput.avg.pct_of_peak_sustained_active"),
                     "l1_global_load_requests": val(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
                     "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](sys.argv[2])
