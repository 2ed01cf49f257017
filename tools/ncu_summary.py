"""Dev helper: text summary of an .ncu-rep (key raw metrics + hottest SASS instructions by stall samples).
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# ncu summary of {rep}\n")
for vals in rows[2:]:
    print("## launch")
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"{h:90s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
si, so = hdr.index("# Samples"), hdr.index("Source")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
agg = {}
for r in data:
    for c in stall:
        agg[hdr[c]] = agg.get(hdr[c], 0) + (int(r[c]) if r[c].isdigit() else 0)
print(f"\n## warp-stall samples: total {tot}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"{k:28s} {v:8d}  {100.0 * v / tot:5.1f}%")
print("\n## hottest SASS instructions (index, samples, instruction, top-2 stall reasons)")
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:28]):
    r = data[i]
    st = sorted(((int(r[c]) if r[c].isdigit() else 0, hdr[c]) for c in stall), reverse=True)[:2]
    print(f"{i:5d} {int(r[si]):7d}  {r[so][:64]:64s} {st}")
