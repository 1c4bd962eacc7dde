"""Hot SASS instructions of one kernel in an ncu report (stall samples, top reasons, shared-memory wavefronts).

    python tools/ncu_hot.py report.ncu-rep kernel-regex [top]
"""
import csv
import io
import subprocess
import sys


def load(rep, regex):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{regex}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = max(i for i, r in enumerate(rows) if r and r[0] == "Address")       # last launch
    hdr = rows[hi]
    out = []
    for r in rows[hi + 1:]:
        if len(r) != len(hdr) or r[0] == "Address":
            break
        out.append(dict(zip(hdr, r)))
    return hdr, out


def num(v):
    try:
        return int(float(v))
    except ValueError:
        return 0


def main():
    rep, regex = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    hdr, out = load(rep, regex)
    for i, o in enumerate(out):
        o["i"] = i
    tot = sum(num(o["# Samples"]) for o in out)
    wf = sum(num(o["L1 Wavefronts Shared"]) for o in out)
    wfi = sum(num(o["L1 Wavefronts Shared Ideal"]) for o in out)
    print(f"instructions {len(out)}  samples {tot}  shared wavefronts {wf} (ideal {wfi})")
    keys = [k for k in hdr if k.startswith("stall_") and "Not" not in k]
    for o in sorted(out, key=lambda o: -num(o["# Samples"]))[:top]:
        st = sorted(((num(o[k]), k[6:]) for k in keys), reverse=True)[:2]
        print(f"{o['i']:5d} {num(o['# Samples']):7d} {num(o['Instructions Executed']):9d} "
              f"wf {num(o['L1 Wavefronts Shared']):9d}/{num(o['L1 Wavefronts Shared Ideal']):9d}  {o['Source'][:64]:64s} {st}")


if __name__ == "__main__":
    main()
