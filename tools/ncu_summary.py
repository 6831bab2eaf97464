"""Key metrics of every launch in an ncu report (raw page), one column per launch."""
import csv
import subprocess
import sys

rep = sys.argv[1]
cols = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else None      # optional: launches to keep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
# every shared-memory / L1 data-bank metric the report holds (the halo kernel's binding resource)
want += [h for h in hdr if ("mem_shared" in h or "data_bank" in h) and "pct_of_peak_sustained_elapsed" in h and h not in want]
for w in want:
    if w not in hdr:
        continue
    i = hdr.index(w)
    vals = [r[i][:60] for r in rows[2:]]
    if cols is not None:
        vals = [vals[c] for c in cols]
    print("%-82s %s  [%s]" % (w, " | ".join(vals), units[i]))
