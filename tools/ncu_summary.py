#!/usr/bin/env python
"""Selected columns of `ncu --set full` captures as one small CSV (first row metric names, second row units -- base units: ns, byte) that
bench.py parses for `roofline.traffic` / `roofline.smem_frac` and that is committed under profiles/.

    python tools/ncu_summary.py profiles/r02_ncu_full_summary.csv gpurun_out/a.ncu-rep gpurun_out/b.ncu-rep ...
"""
import csv
import io
import subprocess
import sys

COLS = [
	"Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
	"dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
	"sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
	"smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
	"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
	"smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def main():
	out, reps = sys.argv[1], sys.argv[2:]
	names, units, rows = None, None, []
	for rep in reps:
		txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
		data = list(csv.reader(io.StringIO(txt)))
		if len(data) < 3:
			continue
		hdr, un = data[0], data[1]
		idx = [hdr.index(c) if c in hdr else -1 for c in COLS]
		if names is None:
			names, units = COLS + ["capture"], [un[i] if i >= 0 else "" for i in idx] + [""]
		for r in data[2:]:
			rows.append([r[i] if i >= 0 else "" for i in idx] + [rep.split("/")[-1]])
	with open(out, "w", newline="") as fh:
		w = csv.writer(fh)
		w.writerow(names); w.writerow(units); w.writerows(rows)
	print(f"{len(rows)} kernels -> {out}")


if __name__ == "__main__":
	main()
