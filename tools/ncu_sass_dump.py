"""Split `ncu --page source --csv` (SASS view) into one annotated listing per kernel:
address, instructions executed, stall samples, opcode text, top-2 stall reasons.
    python tools/ncu_sass_dump.py <sass_all.csv> <kernel substring> <out.txt> [occurrence]"""
import csv
import sys

path, kern, out = sys.argv[1], sys.argv[2], sys.argv[3]
occ = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
match = [i for i in starts if kern in rows[i][1]]
start = match[occ]
end = min([i for i in starts if i > start] + [len(rows)])
hdr = rows[start + 1]
ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot_i = tot_s = 0
with open(out, "w") as f:
    for r in rows[start + 2:end]:
        if len(r) < 10:
            continue
        n = int(r[ii]) if r[ii].isdigit() else 0
        sm = int(r[isamp]) if r[isamp].isdigit() else 0
        tot_i += n
        tot_s += sm
        st = sorted(((int(r[j]), h[6:]) for j, h in stall_cols if r[j].isdigit() and int(r[j]) > 0), reverse=True)[:2]
        f.write(f"{r[0][-5:]} {n:>9} {sm:>6}  {r[1].strip():<72} {st}\n")
print(f"{rows[start][1]}: {tot_i} warp instructions, {tot_s} samples -> {out}")
