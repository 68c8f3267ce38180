"""Summarise an ncu report of gemm_tf32_persistent by warp role: sampled stalls per SASS region (debug aid)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:100])
rows = rows[2:]
tot = sum(int(r[2]) for r in rows)
print("total samples", tot, "lines", len(rows))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i][2]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(idx):
    print(i, rows[i][1].strip()[:64], "samp", rows[i][2], "exec", rows[i][5])
print("-- markers")
for i, r in enumerate(rows):
    t = r[1]
    if any(k in t for k in ("PHASECHK", "LDTM", "UTCBAR", "ARRIVES", "MEMBAR", "BAR.SYNC", "UTCHMMA", "UTCMMA")) and int(r[5]) > 0:
        print(i, t.strip()[:64], r[2], r[5])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, v = rr[0], rr[2]
for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"):
    if k in h:
        print(k, v[h.index(k)], rr[1][h.index(k)])
