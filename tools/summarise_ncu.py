"""Turn ncu CSV output into the small, tracked summaries under profiles/.

    python tools/summarise_ncu.py launches gpurun_out/launches_step.csv profiles/r1_launches_step.md
    python tools/summarise_ncu.py raw gpurun_out/full_X.raw.csv profiles/r1_full_X.md

`launches`: the per-launch `gpu__time_duration.sum` list (B200_PROFILING.md) -> one row per kernel
name with launch count, total and mean device time and its share of the step, own kernels
(libi2p_b200.so) marked.  `raw`: the `--page raw --csv` export of a `--set full` capture -> the
handful of metrics the roofline discussion needs, per captured launch.
"""
import csv
import io
import re
import sys
from collections import OrderedDict

OWN = ("i2p::", "tc::", "conv::")


def _rows(path):
    with open(path, newline="") as fh:
        text = fh.read()
    start = text.find('"ID"')
    if start < 0:
        raise SystemExit("no ncu CSV header in %s" % path)
    return list(csv.DictReader(io.StringIO(text[start:])))


def _short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("at::native::", "at::").replace("(anonymous namespace)::", "")
    return name if len(name) <= 110 else name[:107] + "..."


def launches(src, dst):
    rows = [r for r in _rows(src) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = _short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    own_us = sum(v[1] for k, v in agg.items() if any(o in k for o in OWN))
    with open(dst, "w") as out:
        out.write("# ncu launch list summary: %s\n\n" % src)
        out.write("%d launches, %.1f us serialised device time (cold-cache per-launch times: compare shares); "
                  "own kernels (libi2p_b200.so) %.1f us = %.1f %%\n\n" % (len(rows), total, own_us, 100 * own_us / total))
        out.write("| share | total us | launches | mean us | own | kernel |\n|---:|---:|---:|---:|:-:|---|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write("| %.1f %% | %.1f | %d | %.1f | %s | `%s` |\n" % (100 * us / total, us, n, us / n,
                                                                 "x" if any(o in k for o in OWN) else "", k))
    print("wrote", dst)


def modules(src, labels_json, dst):
    """Per-module device time: the launch list of `profile_step.py marked` + its marker labels."""
    import json
    labels = json.load(open(labels_json))
    rows = [r for r in _rows(src) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg, cur, k = OrderedDict(), "before first marker", 0
    total = 0.0
    for r in rows:
        name = r["Kernel Name"]
        if "spin_kernel" in name:
            cur = labels[k] if k < len(labels) else "after last marker"
            k += 1
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        own = any(o in name for o in OWN)
        a = agg.setdefault(cur, [0, 0.0, 0, 0.0])
        a[0] += 1; a[1] += us
        if own:
            a[2] += 1; a[3] += us
        total += us
    with open(dst, "w") as out:
        out.write("# device time per top-level module (ncu launch list + module-boundary markers): %s\n\n" % src)
        out.write("%d markers matched of %d labels; %.1f us serialised\n\n" % (k, len(labels), total))
        out.write("| share | total us | launches | own us | own launches | segment |\n|---:|---:|---:|---:|---:|---|\n")
        for seg, (n, us, no, uso) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write("| %.1f %% | %.1f | %d | %.1f | %d | %s |\n" % (100 * us / total, us, n, uso, no, seg))
    print("wrote", dst)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum"]


def raw(src, dst):
    with open(src, newline="") as fh:
        text = fh.read()
    start = text.find('"ID"')
    rd = csv.reader(io.StringIO(text[start:]))
    header = next(rd)
    units = next(rd)
    cols = {h: i for i, h in enumerate(header)}
    keep = [m for m in WANT if m in cols]
    extra = [h for h in header if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    keep += [h for h in extra if h not in keep]
    with open(dst, "w") as out:
        out.write("# ncu --set full summary: %s\n\n" % src)
        for row in rd:
            if len(row) < len(header):
                continue
            out.write("## launch %s: `%s` grid %s block %s\n\n" % (row[cols["ID"]], _short(row[cols["Kernel Name"]]),
                                                                  row[cols.get("Grid Size", 0)], row[cols.get("Block Size", 0)]))
            out.write("| metric | value | unit |\n|---|---:|---|\n")
            for m in keep:
                out.write("| %s | %s | %s |\n" % (m, row[cols[m]], units[cols[m]]))
            out.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "modules": modules}[sys.argv[1]](*sys.argv[2:])
