"""Reads a chrome trace written by bench.py (GSB_TRACE / GSB_E2E_TRACE) and reports, for the GPU side: busy time per stream,
the union of kernel activity, every interval longer than 30 us in which NO kernel runs together with the kernel before / after
it, and the concurrency profile (share of the busy time with 1, 2, 3, 4+ kernels in flight).
    python scripts/trace_gaps.py trace.json[.gz]"""
import collections
import gzip
import json
import sys

path = sys.argv[1]
d = json.load(gzip.open(path) if path.endswith(".gz") else open(path))
ev = [e for e in d["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ks = sorted((e for e in ev if e["cat"] != "gpu_memcpy"), key=lambda e: e["ts"])
t0 = ks[0]["ts"]
span = max(e["ts"] + e["dur"] for e in ks) - t0
st = collections.defaultdict(lambda: [0, 0.0])
for e in ks:
    st[e.get("tid")][0] += 1
    st[e.get("tid")][1] += e["dur"]
print("span ms", round(span / 1e3, 3))
for k, v in sorted(st.items(), key=lambda kv: -kv[1][1]):
    print("  stream", k, "kernels", v[0], "busy ms", round(v[1] / 1e3, 2))
# gaps
gaps = []
cur_end, last = ks[0]["ts"] + ks[0]["dur"], ks[0]
for e in ks[1:]:
    if e["ts"] > cur_end:
        if e["ts"] - cur_end > 30:
            gaps.append((cur_end - t0, e["ts"] - cur_end, last["name"][:50], e["name"][:50]))
    if e["ts"] + e["dur"] > cur_end:
        cur_end, last = e["ts"] + e["dur"], e
print("gaps > 30 us:", len(gaps), "total ms", round(sum(g[1] for g in gaps) / 1e3, 3))
for g in gaps[:60]:
    print("  at %8.3f ms  %6.0f us   after %-50s before %s" % (g[0] / 1e3, g[1], g[2], g[3]))
# concurrency profile
pts = []
for e in ks:
    pts.append((e["ts"], 1))
    pts.append((e["ts"] + e["dur"], -1))
pts.sort()
prof, level, prev = collections.defaultdict(float), 0, pts[0][0]
for t, dlt in pts:
    prof[level] += t - prev
    prev = t
    level += dlt
tot = sum(v for k, v in prof.items())
print("kernels in flight: " + ", ".join("%d: %.1f %%" % (k, 100 * v / tot) for k, v in sorted(prof.items())))
