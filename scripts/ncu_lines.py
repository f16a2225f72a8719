"""Aggregate an `ncu --import-source on` report by CUDA source line: executed instructions, stall samples, barrier stalls.
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_lines.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ci = {}
        for i, h in enumerate(hdr):
            ci.setdefault(h, i)
        continue
    if hdr is None or len(r) < 10 or r[0] == "":
        continue
    try:
        ie = float(r[ci["Instructions Executed"]]); smp = float(r[ci["# Samples"]])
    except Exception:
        continue
    def g(name):
        try:
            return float(r[ci[name]])
        except Exception:
            return 0.0
    lines.append((ie, smp, cur, r[0], r[1][:100], g("stall_barrier"), g("stall_long_sb"), g("stall_wait"), g("stall_short_sb"), g("stall_math"), g("stall_no_inst")))
tot_i = sum(l[0] for l in lines); tot_s = sum(l[1] for l in lines)
byfile = collections.Counter(); byfile_s = collections.Counter()
for l in lines:
    byfile[l[2]] += l[0]; byfile_s[l[2]] += l[1]
print("total warp-instructions %.4g, samples %d" % (tot_i, tot_s))
for f in byfile:
    print("  %-18s inst %5.1f%%  samples %5.1f%%" % (f, 100 * byfile[f] / tot_i, 100 * byfile_s[f] / tot_s))
print("--- top lines by samples")
for l in sorted(lines, key=lambda l: -l[1])[:top]:
    print("%5.2f%%s %5.2f%%i bar%5d lsb%5d wait%5d ssb%5d math%5d noi%4d | %s:%s | %s" % (100 * l[1] / tot_s, 100 * l[0] / tot_i, l[5], l[6], l[7], l[8], l[9], l[10], l[2], l[3], l[4]))
print("--- top lines by instructions")
for l in sorted(lines, key=lambda l: -l[0])[:top]:
    print("%5.2f%%i %5.2f%%s | %s:%s | %s" % (100 * l[0] / tot_i, 100 * l[1] / tot_s, l[2], l[3], l[4]))
