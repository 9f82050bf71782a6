"""Top source lines by warp-stall samples for one kernel of an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(out) if l.startswith('"Line No"')][0]
rows = list(csv.reader(out[start:]))
hdr = rows[0]
iS, iL, iSrc, iInst = hdr.index("# Samples"), hdr.index("Line No"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
tot = 0; data = []
for r in rows[1:]:
    if len(r) <= iS or not r[iL]: continue
    try: s = int(r[iS])
    except ValueError: continue
    tot += s
    top = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for i, h in stall_cols), reverse=True)[:2]
    data.append((s, r[iL], r[iSrc].strip()[:95], r[iInst], top))
print("total samples", tot)
for s, ln, src, inst, top in sorted(data, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
    print(f"{100*s/max(tot,1):5.1f}%  L{ln:>4s} inst={inst:>9s} {', '.join(f'{h[6:]}={v}' for v,h in top if v)}  | {src}")
