"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --print-source sass) with nvdisasm -g -c of the
locally built object, to attribute executed instructions / stall samples to device functions and source lines.

usage: python tools/sass_profile.py <sass.csv> <kernel.sass> <mangled-substring> [--lines N]
"""
import csv
import re
import sys
from collections import defaultdict


def parse_nvdisasm(path, kernel_sub):
    ins = []  # (func, file, line, text)
    in_sec = False
    func = "kernel"
    cur = ("?", 0)
    for ln in open(path, errors="replace"):
        if ln.startswith("//---------------------"):
            in_sec = (".text." in ln) and (kernel_sub in ln)
            func = "kernel"
            continue
        if not in_sec:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"^(\$?[\w$.]+):", ln)
        if m:
            lab = m.group(1)
            if lab.startswith("$") and "$" in lab[1:]:
                func = lab.split("$")[-1] if not lab.split("$")[-1].isdigit() else lab
            elif lab.startswith(".text."):
                func = "kernel"
            continue
        m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((func, cur[0], cur[1], m.group(2).strip()))
    return ins


def main():
    sass_csv, nvd, sub = sys.argv[1:4]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    data = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
    ins = parse_nvdisasm(nvd, sub)
    print(f"ncu instructions {len(data)}, nvdisasm instructions {len(ins)}")
    n = min(len(data), len(ins))
    mism = sum(1 for i in range(n) if data[i]["Source"].split()[0:1] != ins[i][3].split()[0:1] and
               data[i]["Source"].split()[1:2] != ins[i][3].split()[0:1])
    print("opcode mismatches:", mism)
    tot_i = sum(int(d["Instructions Executed"]) for d in data)
    tot_s = sum(int(d["# Samples"]) for d in data)
    byf = defaultdict(lambda: [0, 0, 0, 0])
    byl = defaultdict(lambda: [0, 0])
    byop = defaultdict(int)
    for i in range(n):
        d = data[i]
        f = ins[i]
        e, s = int(d["Instructions Executed"]), int(d["# Samples"])
        b = byf[f[0]]
        b[0] += e; b[1] += s
        b[2] += int(d["L1 Wavefronts Shared Excessive"] or 0); b[3] += int(d["L1 Wavefronts Shared"] or 0)
        byl[(f[0][:40], f[1], f[2])][0] += e
        byl[(f[0][:40], f[1], f[2])][1] += s
        op = d["Source"].split()
        op = op[1] if op[0].startswith("@") else op[0]
        byop[op.split(".")[0]] += e
    print(f"\ntotal warp instructions {tot_i:.3e}, samples {tot_s}")
    print("\n== by function: inst%  samples%  smem wavefronts (excess/total)")
    for k, v in sorted(byf.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[0] / tot_i * 100:6.2f} {v[1] / tot_s * 100:6.2f}  {v[2]:>12d}/{v[3]:<12d} {k[:110]}")
    print("\n== by opcode: inst%")
    for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:25]:
        print(f"{v / tot_i * 100:6.2f} {k}")
    print("\n== top source lines: inst% samples%")
    for k, v in sorted(byl.items(), key=lambda kv: -kv[1][1])[:nlines]:
        print(f"{v[0] / tot_i * 100:6.2f} {v[1] / tot_s * 100:6.2f}  {k[1]}:{k[2]}  [{k[0]}]")


if __name__ == "__main__":
    main()
