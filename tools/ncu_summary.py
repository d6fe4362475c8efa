"""Summarise `.ncu-rep` captures (ncu --set full) into the JSON kept under profiles/: per kernel launch the duration,
DRAM bytes, tensor-pipe / issue / warp activity, L2 traffic, registers.  Runs here (no GPU):
    python tools/ncu_summary.py out.json "note" a.ncu-rep [b.ncu-rep ...]"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sectors.sum": "l2_sectors",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "launch__registers_per_thread": "regs",
    "smsp__inst_executed.sum": "warp_inst",
    "sm__cycles_elapsed.max": "cycles",
    "launch__block_size": "block",
    "launch__grid_size": "grid",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0}

out_path, note, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
kernels = []
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        k = {"kernel": r[hdr.index("Kernel Name")], "capture": rep.split("/")[-1]}
        for i, h in enumerate(hdr):
            if h in WANT and r[i] not in ("", "n/a"):
                v = float(r[i].replace(",", ""))
                u = units[i]
                if u in SCALE:
                    v *= SCALE[u]
                elif u in ("us", "usecond"):
                    v, u = v * 1e-3, "ms"
                elif u in ("ns", "nsecond"):
                    v, u = v * 1e-6, "ms"
                k[WANT[h]] = v
                if WANT[h] == "duration":
                    k["duration_unit"] = "ms" if u in ("ms", "msecond") else u
        kernels.append(k)
json.dump({"source": note, "kernels": kernels}, open(out_path, "w"), indent=1)
for k in kernels:
    print(f"{k['kernel'][:70]:70s} {k.get('duration', 0):8.4f} {k.get('duration_unit', '')} tensor {k.get('tensor_pipe_active_pct', 0):5.1f}% "
          f"issue {k.get('issue_active_pct', 0):5.1f}% dram R/W {k.get('dram_read_bytes', 0) / 1e6:8.1f}/{k.get('dram_write_bytes', 0) / 1e6:8.1f} MB")
