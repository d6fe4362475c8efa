"""Summarise `ncu --page source --csv --print-source cuda,sass` output: per source line and per SASS opcode
instruction counts and stall samples of one kernel.
    python tools/ncu_src_summary.py <csv> <kernel substring> [top N]"""
import collections
import csv
import os
import sys

path, kern = sys.argv[1], sys.argv[2]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur_file = cur_fn = None
hdr = None
lines = collections.defaultdict(lambda: [0, 0, ""])  # (file,line) -> [inst, samples, text]
ops = collections.defaultdict(lambda: [0, 0])
stall_tot = collections.Counter()
tot_inst = tot_samp = 0
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = os.path.basename(row[1]); continue
    if row[0] == "Function Name":
        cur_fn = row[1]; continue
    if row[0] == "Line No":
        hdr = row; continue
    if cur_fn is None or kern not in cur_fn or hdr is None or len(row) < len(hdr) - 2:
        continue
    i_inst = hdr.index("Instructions Executed")
    i_samp = hdr.index("# Samples")
    if row[0]:  # source line row
        key = (cur_file, int(row[0]))
        lines[key][0] += num(row[i_inst])
        lines[key][1] += num(row[i_samp])
        lines[key][2] = row[1].strip()[:90]
    else:  # SASS row
        op = row[3].split()[0] if not row[3].strip().startswith("@") else row[3].split()[1]
        n, s = num(row[i_inst]), num(row[i_samp])
        ops[op][0] += n; ops[op][1] += s
        tot_inst += n; tot_samp += s
        for j, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h and row[j] not in ("", "-"):
                stall_tot[h] += num(row[j])
print(f"kernel ~ {kern}: {tot_inst} warp instructions, {tot_samp} samples")
print("stalls:", ", ".join(f"{k[6:]} {v*100/max(1,tot_samp):.1f}%" for k, v in stall_tot.most_common(10)))
print("-- by source line (inst share, sample share)")
for (f, l), (n, s, t) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n*100/tot_inst:5.1f}% {s*100/max(1,tot_samp):5.1f}%  {f}:{l}  {t}")
print("-- by opcode")
for op, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n*100/tot_inst:5.1f}% {s*100/max(1,tot_samp):5.1f}%  {op}")
