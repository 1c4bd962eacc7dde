"""gpurun_out/f (written by tools/run_profiles.sh on the GPU box) -> profiles/rN_* (tracked).

    python tools/collect_profiles.py [round]
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "f")
DST = os.path.join(ROOT, "profiles")
R = "r" + (sys.argv[1] if len(sys.argv) > 1 else "2")


def copy_json(src, dst):
    line = [l for l in open(os.path.join(SRC, src)) if l.startswith("{")][-1]
    json.loads(line)
    open(os.path.join(DST, dst), "w").write(line)


def main():
    copy_json("bench.json", f"{R}_bench_bf16.json")
    copy_json("bench_ref.json", f"{R}_bench_reference_arm.json")
    if os.path.exists(os.path.join(SRC, "bench_fp32.json")):
        copy_json("bench_fp32.json", f"{R}_bench_fp32.json")
    if os.path.exists(os.path.join(SRC, "stream.json")):
        copy_json("stream.json", f"{R}_stream_bench_1gpu.json")
    for src, dst in (("latency.json", f"{R}_latency.json"), ("h2d_matrix.json", f"{R}_h2d_matrix_1gpu.json")):
        if os.path.exists(os.path.join(SRC, src)):
            copy_json(src, dst)
    shutil.copy(os.path.join(SRC, "smoke.log"), os.path.join(DST, f"{R}_smoke.log"))
    # launch list: keep kernel name, grid, block, duration
    rows = [r for r in csv.reader(open(os.path.join(SRC, "launches.csv"))) if len(r) > 10]
    head = rows[0]
    keep = [head.index(k) for k in ("ID", "Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")]
    with open(os.path.join(DST, f"{R}_launches.csv"), "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in keep])
    lb = os.path.join(SRC, "launches_bench.csv")
    if os.path.exists(lb):
        rows_b = [r for r in csv.reader(open(lb)) if len(r) > 10]
        hb = rows_b[0]
        keep_b = [hb.index(k) for k in ("ID", "Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")]
        with open(os.path.join(DST, f"{R}_launches_bench.csv"), "w", newline="") as f:
            w = csv.writer(f)
            for r in rows_b:
                w.writerow([r[i] for i in keep_b])
    tool = os.path.join(ROOT, "tools", "ncu_summary.py")
    subprocess.run([sys.executable, tool, os.path.join(SRC, "prof_full.ncu-rep"), os.path.join(DST, f"{R}_ncu_full_summary.csv")], check=True)
    if os.path.exists(os.path.join(SRC, "prof_feat.ncu-rep")):       # absent after tools/run_profiles_short.sh
        subprocess.run([sys.executable, tool, os.path.join(SRC, "prof_feat.ncu-rep"), os.path.join(DST, f"{R}_ncu_feat_kernel_65536.csv"), "feat_kernel"], check=True)
    # per-clip shares from the launch list (second pass = launches 7..13)
    dur = {}
    for r in rows[1:]:
        dur.setdefault(r[head.index("Kernel Name")].split("(")[0], []).append(float(r[head.index("Metric Value")]))
    tot = sum(v[-1] for v in dur.values())
    with open(os.path.join(DST, f"{R}_kernel_shares.txt"), "w") as f:
        f.write("kernel, us per 16384-clip slab (ncu, cold, serialised), ns/clip, share\n")
        for k, v in dur.items():
            f.write(f"{k}, {v[-1] / 1e3:.1f}, {v[-1] / 16384:.2f}, {v[-1] / tot:.3f}\n")
        f.write(f"total, {tot / 1e3:.1f}, {tot / 16384:.2f}, 1.000\n")
    print(open(os.path.join(DST, f"{R}_kernel_shares.txt")).read())


if __name__ == "__main__":
    main()
