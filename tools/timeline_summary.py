"""Summarise a kernel timeline written by `tools/profile_step.py timeline`: concurrency histogram, per-stream busy time
and gaps, the tail of the step.  python tools/timeline_summary.py gpurun_out/timeline.csv [tail_from_us]"""
import collections
import csv
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
for r in rows:
    r["s"] = float(r["start_us"]); r["d"] = float(r["dur_us"]); r["e"] = r["s"] + r["d"]
# the profiler separates the graph's first two fills from the rest by a launch gap: drop everything before the largest gap in the first 2 ms
rows.sort(key=lambda r: r["s"])
t0 = 0.0
for a, b in zip(rows, rows[1:]):
    if b["s"] < 2000 and b["s"] - a["e"] > 300:
        t0 = b["s"]
rows = [r for r in rows if r["s"] >= t0]
span = max(r["e"] for r in rows) - t0
print("step span %.0f us, %d kernels, kernel time %.0f us" % (span, len(rows), sum(r["d"] for r in rows)))
ev = []
for r in rows:
    ev.append((r["s"], 1)); ev.append((r["e"], -1))
ev.sort()
cur, last, hist = 0, t0, collections.Counter()
for t, dv in ev:
    hist[min(cur, 4)] += t - last; last = t; cur += dv
print("time with k kernels in flight:", {k: round(v) for k, v in sorted(hist.items())})
st = collections.defaultdict(list)
for r in rows:
    st[r["stream"]].append(r)
print("stream: busy / first / last / gaps inside")
for k, v in sorted(st.items(), key=lambda kv: -sum(r["d"] for r in kv[1])):
    gaps = sum(max(0.0, b["s"] - a["e"]) for a, b in zip(v, v[1:]))
    print("  %4s busy %6.0f n %3d  %6.0f .. %6.0f  gaps %6.0f" % (k, sum(r["d"] for r in v), len(v), v[0]["s"] - t0, v[-1]["e"] - t0, gaps))
if len(sys.argv) > 2 and sys.argv[2] != "path":
    frm = float(sys.argv[2]) + t0
    for r in rows:
        if r["s"] >= frm and r["d"] > 8:
            print("%7.0f %6.1f %4s %s" % (r["s"] - t0, r["d"], r["stream"], r["name"][:90]))


def critical_path(rows, t0):
    """Walk back from the last kernel: the predecessor of a kernel is the one (any stream) that ended last before it
    started -- what it waited for.  -> time on the path by kernel name, and the idle time between the links."""
    import collections
    by_end = sorted(rows, key=lambda r: r["e"])
    ends = [r["e"] for r in by_end]
    import bisect
    cur = max(rows, key=lambda r: r["e"])
    on_path, gaps, n = collections.Counter(), 0.0, 0
    while True:
        on_path[cur["name"].replace("void ", "").replace("at::native::", "")[:70]] += cur["d"]
        n += 1
        i = bisect.bisect_right(ends, cur["s"] + 0.05) - 1
        while i >= 0 and by_end[i] is cur:
            i -= 1
        if i < 0:
            break
        prev = by_end[i]
        if prev["e"] < t0 or cur["s"] - prev["e"] > 200:
            break
        gaps += max(0.0, cur["s"] - prev["e"])
        cur = prev
    return on_path, gaps, n


if len(sys.argv) > 2 and sys.argv[2] == "path":
    on_path, gaps, n = critical_path(rows, t0)
    print("critical path: %d kernels, %.0f us of kernel time, %.0f us between them" % (n, sum(on_path.values()), gaps))
    for k, v in on_path.most_common(40):
        print("%7.1f %s" % (v, k))
