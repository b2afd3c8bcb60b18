"""Top source lines by instructions executed, per kernel, from an ncu report (needs -lineinfo + --import-source on)."""
import csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
func, hdr, acc = None, None, {}
def flush():
    if func and acc:
        tot = sum(v[0] for v in acc.values()); tots = sum(v[1] for v in acc.values())
        print(f"== {func[:100]}  inst={tot:.3g} samples={tots}")
        for ln, (ins, smp, src) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"  L{ln:>4} {100*ins/max(tot,1):5.1f}% inst {100*smp/max(tots,1):5.1f}% smp  {src.strip()[:110]}")
for r in rows:
    if not r: continue
    if r[0] == "Function Name":
        flush(); func, acc = r[1], {}
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try: acc[int(r[0])] = (float(r[ie] or 0), float(r[sm] or 0), r[1])
        except ValueError: pass
flush()
