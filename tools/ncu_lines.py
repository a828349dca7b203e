"""Per-source-line executed warp instructions from an ncu report (cuda,sass source page).

    python tools/ncu_lines.py report.ncu-rep [top [kernel-regex]]
"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
if len(sys.argv) > 3:
    cmd += ["-k", "regex:" + sys.argv[3]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; lines = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0] != "":   # a CUDA source line (aggregate over its SASS)
        idx = hdr.index("Instructions Executed")
        try: n = float(r[idx])
        except ValueError: n = 0
        sidx = hdr.index("# Samples")
        try: smp = float(r[sidx])
        except ValueError: smp = 0
        key = (cur_file, int(r[0]), r[1].strip())
        a = lines.setdefault(key, [0, 0]); a[0] += n; a[1] += smp
tot = sum(v[0] for v in lines.values()); tots = sum(v[1] for v in lines.values())
print("total warp instructions %.0f, samples %.0f" % (tot, tots))
for (f, ln, src), (n, smp) in sorted(lines.items(), key=lambda kv: -kv[1][1 if "--by-samples" in sys.argv else 0])[:top]:
    print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100 * n / tot, 100 * smp / max(tots, 1), f, ln, src[:110]))
