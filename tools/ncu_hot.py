"""Top stall sites of one kernel from an ncu report's source page (SASS view).
usage: python tools/ncu_hot.py report.ncu-rep <kernel regex> [n]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')
txt = '"Kernel Name"' + blocks[1]
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
iS, iSm, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > iI and r[iSm].isdigit()]
tot_s = sum(int(r[iSm]) for r in body); tot_i = sum(int(r[iI]) for r in body)
print(rows[0][1], "samples", tot_s, "warp-instructions", tot_i, "SASS lines", len(body))
top = sorted(range(len(body)), key=lambda i: -int(body[i][iSm]))[:n]
for i in sorted(top):
    r = body[i]
    print("%5d %6.2f%% smp  %8s exec  %s" % (i, 100.0 * int(r[iSm]) / tot_s, r[iI], r[iS].strip()[:110]))
