"""Per-source-line stall samples / executed instructions of an .ncu-rep (kernel built with -lineinfo, --import-source on).
usage: python tools/ncu_lines.py REPORT.ncu-rep [top_n]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])  # samples, inst, fp64 inst, text
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci, cn = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ci:
        continue
    if r[0] != "":
        cur = (fname, int(r[0]))
        agg[cur][3] = r[1].strip()[:90]
    elif cur is not None and r[2] not in ("", "..."):
        try:
            n, s = int(r[ci]), int(r[cn])
        except ValueError:
            continue
        a = agg[cur]
        a[0] += s
        a[1] += n
        op = r[3].strip().split()
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
        if op.split(".")[0] in ("DFMA", "DADD", "DMUL", "DSETP"):
            a[2] += n
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {ts}, warp instructions {ti}")
print("samples%  inst%  fp64%ofline  file:line  source")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/ts:6.2f} {100*a[1]/ti:6.2f} {100*a[2]/max(a[1],1):6.1f}  {k[0]}:{k[1]}  {a[3]}")
