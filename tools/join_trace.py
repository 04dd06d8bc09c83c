"""Join a DG_TRACE=1 log (one line per tcgen05 GEMM launch) with the ncu launch list of the same run:
   python tools/join_trace.py <launches.csv> <stderr log> [kernel substring]"""
import collections
import re
import sys

sys.path.insert(0, "tools")
from summarize_launches import load

if __name__ == "__main__":
    pat = sys.argv[3] if len(sys.argv) > 3 else "gemm"
    rows = [r for r in load(sys.argv[1]) if pat in r["Kernel Name"]]
    tr = [l.strip() for l in open(sys.argv[2]) if "DG_TRACE gemm" in l][-len(rows):]
    agg = collections.OrderedDict()
    for r, t in zip(rows, tr):
        m = dict(re.findall(r"(\w+)=(\S+)", t))
        key = (m["M"], m["N"], m["K"], m["taps"], m.get("splits", "1"), m["grid"], m["geglu"], m.get("ln", "0"), m.get("res", "0"))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    print("M N K taps splits grid geglu ln res | count avg_us TFLOP/s total_us")
    tot = 0.0
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        M, N, K = int(k[0]), int(k[1]), int(k[2])
        print(*k, "|", c, round(t / c, 1), round(2.0 * M * N * K / (t / c * 1e-6) / 1e12, 1), round(t, 1))
        tot += t
    print("total us", round(tot, 1))
