"""Where a kernel's warps wait: stall samples per SASS instruction from an ncu report (needs --import-source on).

    python tools/ncu_stalls.py gpurun_out/x.ncu-rep <clips> [top]
"""
import collections, csv, io, subprocess, sys
rep, N = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))[2:]
ins = []
for r in rows:
    try:
        ins.append((r[1].strip(), int(r[2]), int(r[5])))
    except (ValueError, IndexError):
        pass
tot = sum(x[1] for x in ins)
print(f"{len(ins)} instructions, {sum(x[2] for x in ins) / N:.0f} executed per clip, {tot} stall samples")
reg, regi = collections.Counter(), collections.Counter()
for s, st, ex in ins:
    k = round(ex / N)
    reg[k] += st
    regi[k] += ex / N
for k, v in reg.most_common(10):
    print(f"  executed x{k:<4d} per clip: stall {100 * v / tot:5.1f}%   {regi[k]:8.0f} instr/clip")
idx = sorted(range(len(ins)), key=lambda i: -ins[i][1])[:top]
for i in sorted(idx):
    s, st, ex = ins[i]
    print(f"{100 * st / tot:5.2f}%  x{ex / N:5.0f}  {s[:72]:72s} | prev: {ins[i - 1][0][:46] if i else ''}")
