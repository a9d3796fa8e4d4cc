#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown file for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r02/frames_allegro.md "title" \
           [--traffic profiles/roofline_traffic.json --frames 65536 --bytes-per-frame 380 --build-id <dexr_build_id>]

With --traffic, also writes the per-launch figures bench.py turns into `roofline.traffic`, `roofline.issue` and `roofline.fp32`
(DRAM bytes, warp instructions, FP32 operations = 2 x FFMA + FADD + FMUL thread instructions), stamped with the library build
id the capture was taken from; bench.py refuses the file when the loaded library reports another id.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__thread_inst_executed.sum",
]


def _opt(name, default=None):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {title}", "", f"source: `{rep}` (`ncu --set full --clock-control none --import-source on`), {len(data)} launch(es) captured; "
             "per-launch values below are from the first captured launch.", ""]
    name = data[0][col["Kernel Name"]] if "Kernel Name" in col else "?"
    lines += [f"kernel: `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in col:
            lines.append(f"| {k} | {data[0][col[k]]} | {units[col[k]]} |")
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(data[0][i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    lines += ["", "warp stall reasons (average warps stalled per issue-active cycle):", "", "| reason | warps |", "|---|---|"]
    for v, n in sorted(stalls, reverse=True):
        if v >= 0.005:
            lines.append(f"| {n} | {v:.3f} |")
    pipes = []
    for h, i in col.items():
        if h.startswith("sm__inst_executed_pipe_") and h.endswith(".sum"):
            try:
                pipes.append((float(data[0][i]), h))
            except ValueError:
                pass
    if pipes:
        lines += ["", "instructions by pipe:", "", "| pipe | warp instructions |", "|---|---|"]
        for v, n in sorted(pipes, reverse=True)[:12]:
            lines.append(f"| {n} | {v:.0f} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    traffic = _opt("--traffic")
    if traffic:
        import json

        def val(k, scale=1.0):
            if k not in col:
                return None
            v = float(data[0][col[k]].replace(",", ""))
            u = units[col[k]]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
            return v * mult * scale

        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        ffma, fadd, fmul = (val(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum") for o in ("ffma", "fadd", "fmul"))
        frames_n, bpf = int(_opt("--frames", "65536")), int(_opt("--bytes-per-frame", "380"))
        rec = {"build_id": _opt("--build-id"), "kernel": name, "frames_per_launch": frames_n,
               "dram_bytes_per_launch": None if rd is None else int(rd + wr), "read": rd, "write": wr,
               "algorithmic_bytes_per_launch": frames_n * bpf, "warp_inst_per_launch": val("smsp__inst_executed.sum"),
               "fp32_flop_per_launch": None if ffma is None else 2 * ffma + (fadd or 0) + (fmul or 0),
               "ffma_thread_inst": ffma, "fadd_thread_inst": fadd, "fmul_thread_inst": fmul,
               "duration_us_under_ncu": val("gpu__time_duration.sum") if "gpu__time_duration.sum" in col else None,
               "source": f"{out} (ncu --set full, first captured launch)"}
        open(traffic, "w").write(json.dumps(rec, indent=1) + "\n")
        print("wrote", traffic)


if __name__ == "__main__":
    main()
