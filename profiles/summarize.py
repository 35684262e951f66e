"""Summarise ncu outputs into small text files that can be committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/<name>_launches.txt
    python profiles/summarize.py raw gpurun_out/prof.ncu-rep       > profiles/<name>_raw.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    text = open(path, errors="replace").read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("satfill::", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += us
        a[2] = max(a[2], us)
        total += us
    print(f"# {path}: {len(rows)} launches, {total / 1e3:.3f} ms of kernel time (ncu: cold cache, serialised -> compare SHARES)")
    print(f"{'kernel':60s} {'launches':>9s} {'total_ms':>10s} {'share':>7s} {'avg_us':>9s} {'max_us':>9s}")
    for name, (n, us, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {n:9d} {us / 1e3:10.3f} {us / total:7.1%} {us / n:9.1f} {mx:9.1f}")


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "dram__cycles_elapsed.avg.per_second", "gpc__cycles_elapsed.avg.per_second",
    # where the reads are served from (do the frame halos hit in L2? at which granularity is
    # DRAM filled?)
    "dram__sectors_read.sum", "dram__sectors_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "lts__t_bytes.sum", "smsp__inst_executed.sum",
]  # fmt: skip


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path} (ncu --set full --clock-control none)")
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        rd = wr = None
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:70s} {r[i]:>14s} {units[i]}")
                if w == "dram__bytes_read.sum":
                    rd = (float(r[i].replace(",", "")), units[i])
                if w == "dram__bytes_write.sum":
                    wr = (float(r[i].replace(",", "")), units[i])
        if rd and wr:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]
            i = hdr.index("gpu__time_duration.sum")
            t = float(r[i].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[units[i]]
            print(f"  {'traffic = dram read + write':70s} {tot / 1e9:14.4f} GB  -> {tot / t / 1e9:.0f} GB/s under ncu")


def traffic(path, units):
    """dram bytes per unit of work (unknown x band) of the captured launch, as JSON: bench.py scales it to its own launches.
        python profiles/summarize.py traffic rep.ncu-rep <unknown-bands the captured launch processed>"""
    import json

    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units_row, r = rows[0], rows[1], rows[2]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for w in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(w)
        tot += float(r[i].replace(",", "")) * scale[units_row[i]]
    print(json.dumps({"kernel": re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", ""),
                      "dram_bytes": tot, "units": float(units), "dram_bytes_per_unit": tot / float(units), "capture": path}))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3])
    else:
        {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
