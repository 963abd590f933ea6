#!/usr/bin/env python
"""Summarise an .ncu-rep: the metrics the profiles/ tables quote, and the executed opcode mix.

    python tools/ncu_summary.py report.ncu-rep [out_prefix] [--hw hw_counters.json WORKLOAD RAYS WARP_RHS [traffic.csv TRAFFIC_RAYS STEPS]]

Writes <prefix>_ncu_summary.txt and <prefix>_opcode_mix.txt (needs `ncu` on PATH; no GPU).  With --hw it also
records the hardware-counted figures bench.py quotes beside the algorithmic roofline (roofline.hw) under WORKLOAD
in hw_counters.json: FP64 instructions per cycle, pipe utilisations, instructions per RHS (executed warp
instructions / WARP_RHS of the captured launch) and, if given, the DRAM bytes of a full-size launch."""
import collections, csv, io, json, os, re, subprocess, sys

METRICS = """dram__bytes_read.sum dram__bytes_write.sum gpu__time_duration.sum l1tex__t_sector_hit_rate.pct
launch__registers_per_thread lts__t_sector_hit_rate.pct sm__cycles_elapsed.avg sm__cycles_elapsed.avg.per_second
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__issue_active.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active dram__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed
lts__t_sectors.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed
lts__throughput.avg.pct_of_peak_sustained_elapsed
smsp__inst_executed.sum smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed""".split()


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, vals = rows[0], rows[1], rows[2]
    return {n: (u, v) for n, u, v in zip(names, units, vals)}, vals[4]


def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next((i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" == c for c in r)), None)
    if hdr is None:
        return None
    h = rows[hdr]
    i_src, i_exec = h.index("Source"), h.index("Instructions Executed")
    i_samp = h.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in h else None
    mix, samples, static = collections.Counter(), collections.Counter(), 0
    for r in rows[hdr + 1:]:
        if len(r) <= i_exec or not r[i_src].strip():
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[i_src])
        if not m:
            continue
        op = m.group(2)
        static += 1
        mix[op] += int(float(r[i_exec] or 0))
        if i_samp is not None:
            samples[op] += int(float(r[i_samp] or 0))
    return static, mix, samples


def metric(vals, name):
    key = next((k for k in vals if k == name or k.endswith("." + name)), None)
    return float(vals[key][1].replace(",", "")) if key and vals[key][1] != "" else None


def template_flag(kernel, i):
    """trace_kernel<BK, CK, MATH, UNI, DMAP, SG>: parameter i as a bool (None if the name has no such parameter)"""
    m = re.search(r"<([^>]*)>", kernel)
    p = [t.strip() for t in m.group(1).split(",")] if m else []
    return (p[i] in ("1", "true")) if i < len(p) else None


def record_hw(vals, kernel, rep, args):
    path, workload, rays, warp_rhs = args[0], args[1], int(args[2]), float(args[3])
    db = {}
    if os.path.exists(path):
        db = json.load(open(path))
    inst = metric(vals, "smsp__inst_executed.sum")
    e = {
        "kernel": kernel, "capture": os.path.basename(rep) + " (ncu --set full --clock-control none), summarised in profiles/",
        "rays": rays, "deep_map": template_flag(kernel, 4), "same_grid": template_flag(kernel, 5),
        "pipe_fp64_pct": metric(vals, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "dadd_per_cycle": metric(vals, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"),
        "dmul_per_cycle": metric(vals, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"),
        "dfma_per_cycle": metric(vals, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"),
        "issue_active_pct": metric(vals, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        "l1_lsu_wavefronts_pct": metric(vals, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "registers": metric(vals, "launch__registers_per_thread"),
        "instr_per_rhs": inst / warp_rhs if inst and warp_rhs else None,
        "gpu_time_ms": metric(vals, "gpu__time_duration.sum"),
    }
    if len(args) >= 7:
        rows = list(csv.reader(open(args[4])))
        get = lambda name: next(float(r[-1].replace(",", "")) for r in rows if len(r) > 3 and r[-3] == name)
        e.update(traffic_rays=int(args[5]), rk4_steps=int(args[6]), dram_bytes_read=get("dram__bytes_read.sum"),
                 dram_bytes_write=get("dram__bytes_write.sum"), traffic_capture=os.path.basename(args[4]))
    db[workload] = e
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)
    print("hw counters of", workload, "->", path)


def main():
    rep = sys.argv[1]
    hw = None
    if "--hw" in sys.argv:
        i = sys.argv.index("--hw")
        hw = sys.argv[i + 1:]
        del sys.argv[i:]
    prefix = sys.argv[2] if len(sys.argv) > 2 else rep.rsplit(".", 1)[0]
    vals, kernel = raw(rep)
    if hw:
        record_hw(vals, kernel, rep, hw)
    with open(prefix + "_ncu_summary.txt", "w") as f:
        f.write(f"# ncu -i {rep.split('/')[-1]} --page raw --csv   (kernel: {kernel})\n")
        stalls = sorted(n for n in vals if "issue_stalled" in n and n.endswith("per_issue_active.ratio") and "not_issued" not in n)
        for n in METRICS + stalls:
            key = next((k for k in vals if k == n or k.endswith("." + n)), None)
            if key and vals[key][1] != "":
                f.write(f"{n:<88} {vals[key][0]:<16} {vals[key][1]}\n")
    src = source(rep)
    if src:
        static, mix, samples = src
        tot, stot = sum(mix.values()), max(sum(samples.values()), 1)
        with open(prefix + "_opcode_mix.txt", "w") as f:
            f.write(f"static instrs {static} executed warp-instrs {tot} samples {sum(samples.values())}\n")
            for op, c in mix.most_common(40):
                f.write(f"{op:<28} {c:>14} {100 * c / tot:6.2f}%  samples {100 * samples[op] / stot:6.2f}%\n")


if __name__ == "__main__":
    main()
