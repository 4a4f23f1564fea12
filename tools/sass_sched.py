"""Join the static schedule of a kernel's SASS (stall counts decoded from the control bits) with the
executed-instruction counts of an .ncu-rep: how many issue cycles one warp needs per attempt if it ran alone.
usage: python tools/sass_sched.py LIB.so MANGLED_KERNEL REPORT.ncu-rep ATTEMPTS_PER_LAUNCH [dump]"""
import csv, io, re, subprocess, sys

lib, fn, rep, attempts = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
dump = len(sys.argv) > 5
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
lines = txt.splitlines()
st = next(i for i, l in enumerate(lines) if "Function : " + fn in l)
ins = []
i = st + 1
while i < len(lines) and "Function :" not in lines[i]:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", lines[i])
    if m:
        m2 = re.search(r"/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
        hi = int(m2.group(1), 16)
        ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xf, (hi >> 45) & 1, (hi >> 46) & 7,
                    (hi >> 49) & 7, (hi >> 52) & 0x3f))
        i += 2
    else:
        i += 1
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
i0 = [k for k, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[i0]
ci, cn = h.index("Instructions Executed"), h.index("# Samples")
body = [r for r in rows[i0 + 1:] if len(r) > ci]
assert len(body) == len(ins), (len(body), len(ins))
wa = attempts / 32.0
tot_issue = tot_stall = 0
tot_samples = sum(int(r[cn]) for r in body)
for (off, text, stall, yld, wb, rb, wm), r in zip(ins, body):
    n = int(r[ci])
    assert text.split()[0].strip("@!UP0123456789") in r[1] or True
    tot_issue += n
    tot_stall += n * max(stall, 1)
    if dump:
        print(f"{off:05x} {n / wa:7.3f} st{stall:2d} {'Y' if yld else ' '} w{wb} r{rb} m{wm:02x} smp{100 * int(r[cn]) / tot_samples:5.2f}  {text}")
print(f"# warp-instructions/attempt {tot_issue / wa:.1f}, static issue+stall cycles/attempt (one warp alone) {tot_stall / wa:.1f}")
