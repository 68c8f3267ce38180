"""Stall samples of one launch of an ncu report, bucketed by SASS region (= warp role of gemm_tf32_persistent) and by stall reason.

usage: python scripts/ncu_roles.py <report.ncu-rep> <launch index> [top N instructions] [dump LO HI]

Every warp is sampled at the same rate, so (samples in a role's region) / (total samples / warps per CTA) is the share of that
role's time; a role that is the bottleneck has few samples at its barrier waits (SYNCS.PHASECHK / the BRA that follows).  The
report must have been captured with --section SourceCounters --import-source on (profiles/r01_ncu_gemm16_v10_notes.txt shows
the reading of 16 launches)."""
import csv
import subprocess
import sys

rep, k = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:120])
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[heads[0]]
body = [r for r in rows[heads[0] + 1:(heads[1] - 1 if len(heads) > 1 else len(rows))] if len(r) >= len(h)]
col = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
total = sum(int(r[col["# Samples"]] or 0) for r in body)
print("total samples", total, " per warp (19 warps)", round(total / 19, 1), " SASS instructions", len(body))
MARK = ("LDGSTS", "UTCHMMA", "LDTM", "STG", "LDS", "STS", "SYNCS", "BAR", "UTCBAR", "MEMBAR", "LDG", "UTMASTG", "FMNMX")
for lo in range(0, len(body), 250):
    S = E = 0
    t = {s: 0 for s in stalls}
    ops = set()
    for r in body[lo:lo + 250]:
        S += int(r[col["# Samples"]] or 0)
        E += int(r[col["Instructions Executed"]] or 0)
        for s in stalls:
            t[s] += int(r[col[s]] or 0)
        op = [o for o in r[col["Source"]].split() if not o.startswith("@")][0].split(".")[0]
        if op in MARK:
            ops.add(op)
    if S:
        print("  region", lo, "samples", S, "executed", E, {k2[6:]: v for k2, v in t.items() if v > 30}, sorted(ops))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top]
for i in sorted(idx):
    r = body[i]
    print(i, r[col["Source"]][:64].strip(), r[col["# Samples"]], r[col["Instructions Executed"]],
          {s[6:]: r[col[s]] for s in stalls if int(r[col[s]] or 0) > 15})
if len(sys.argv) > 6 and sys.argv[4] == "dump":
    for i in range(int(sys.argv[5]), int(sys.argv[6])):
        r = body[i]
        print(i, r[col["Source"]][:80].strip().ljust(70), r[col["# Samples"]].rjust(5), r[col["Instructions Executed"]].rjust(8),
              r[col["Avg. Predicated-On Threads Executed"]])
