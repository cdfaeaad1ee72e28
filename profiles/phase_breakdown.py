"""Per-phase instruction / stall-sample shares of halfspace_kernel from an .ncu-rep with -lineinfo source (run here, no GPU).
   python profiles/phase_breakdown.py gpurun_out/x.ncu-rep <halfspaces in the launch> [groups.json]
Groups are (name, first line, last line) ranges of csrc/halfspace_kernel.cuh; lines of inlined helpers are reported per line."""
import csv
import subprocess
import sys
import json

rep, B = sys.argv[1], int(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = None
per = {}   # (file, line) -> [inst, samples, stall dict]
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "" or not r[0].isdigit():
        continue
    ix, sx = hdr.index("Instructions Executed"), hdr.index("# Samples")
    try:
        n, s = int(r[ix]), int(r[sx])
    except ValueError:
        continue
    st = {}
    for j, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                st[h] = int(r[j])
            except ValueError:
                pass
    key = (cur_file, int(r[0]))
    e = per.setdefault(key, [0, 0, {}, r[1]])
    e[0] += n
    e[1] += s
    for k, v in st.items():
        e[2][k] = e[2].get(k, 0) + v
tot = sum(e[0] for e in per.values())
totS = sum(e[1] for e in per.values())
print(f"total warp-instructions {tot}  per halfspace {tot / B:.0f}; stall samples {totS}")
groups = json.load(open(sys.argv[3])) if len(sys.argv) > 3 else None
if groups:
    for name, f, lo, hi in groups:
        sel = [e for (ff, l), e in per.items() if ff == f and lo <= l <= hi]
        n, s = sum(e[0] for e in sel), sum(e[1] for e in sel)
        st = {}
        for e in sel:
            for k, v in e[2].items():
                st[k] = st.get(k, 0) + v
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        print(f"{name:22s} inst/hs {n / B:8.1f} ({100 * n / tot:5.1f}%)  samples {100 * s / totS:5.1f}%  " +
              " ".join(f"{k[6:]}={100 * v / max(s, 1):.0f}%" for k, v in top))
print("top lines by instructions:")
for (f, l), e in sorted(per.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"  {f}:{l:5d} inst/hs {e[0] / B:7.1f} samples {100 * e[1] / totS:4.1f}%  {e[3].strip()[:110]}")
