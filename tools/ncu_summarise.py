"""Read an `ncu --set full` report (brought back in gpurun_out/) and write the artefacts committed under profiles/:

  python tools/ncu_summarise.py gpurun_out/prof.ncu-rep profiles/r1m

  -> profiles/r1m_ncu_summary.csv   one row per metric, one column per captured kernel launch (the raw page, transposed)
     profiles/r1m_ncu_traffic.json  DRAM bytes read / written and duration per kernel (bench.py: roofline.traffic)
     profiles/r1m_ncu_source_mix.txt per-opcode executed instructions and shared-memory wavefronts of the Riccati kernels
"""
import collections
import csv
import io
import json
import subprocess
import sys


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def source_mix(rep, kernel):
    p = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel],
                       capture_output=True, text=True)
    if p.returncode != 0:
        return None
    rows = list(csv.reader(io.StringIO(p.stdout)))
    hdr = None
    for i, r in enumerate(rows):
        if "Source" in r and any("Instructions Executed" == c for c in r):
            hdr, body = r, rows[i + 1:]
            break
    if hdr is None:
        return None
    ci, cs = hdr.index("Source"), hdr.index("Instructions Executed")
    cw = next((hdr.index(c) for c in hdr if c.startswith("L1 Wavefronts Shared") and "Excessive" not in c and "Ideal" not in c), None)
    cwi = next((hdr.index(c) for c in hdr if c.startswith("L1 Wavefronts Shared Ideal")), None)
    ex, wf, wfi = collections.Counter(), collections.Counter(), collections.Counter()
    for r in body:
        if len(r) <= max(ci, cs):
            continue
        toks = r[ci].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.rstrip(";")
        try:
            ex[op] += int(float(r[cs] or 0))
            if cw is not None:
                wf[op] += int(float(r[cw] or 0))
            if cwi is not None:
                wfi[op] += int(float(r[cwi] or 0))
        except ValueError:
            pass
    return ex, wf, wfi


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    what = sys.argv[3] if len(sys.argv) > 3 else "bench.py --steps 1 --warmup 1, C3 B=16384, shipped configuration"
    names, units, rows = raw_page(rep)
    kcol = names.index("Kernel Name")
    kernels = [r[kcol] for r in rows]
    with open(prefix + "_ncu_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + kernels)
        for j, (nm, un) in enumerate(zip(names, units)):
            if nm in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"):
                continue
            w.writerow([nm, un] + [r[j] for r in rows])

    def val(r, metric):
        j = names.index(metric)
        v, u = float(r[j].replace(",", "")), units[j]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
        return v * scale.get(u, 1.0)

    # per-config traffic files: argv[4] = "c3:4,c5:2,c4:1,c2:1" says how many consecutive captured launches belong to which
    # config (the order tools/ncu_case.py runs them in); the first group keeps the plain name bench.py looks up for C3
    split = sys.argv[4] if len(sys.argv) > 4 else "c3:%d" % len(rows)
    pos = 0
    for k, item in enumerate(split.split(",")):
        cfg, cnt = item.split(":")
        traffic = {"source": "%s_ncu_summary.csv (ncu --set full --clock-control none, %s), config %s" % (prefix, what, cfg)}
        for r in rows[pos:pos + int(cnt)]:
            traffic[r[kcol]] = {"dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
                                "gpu_time_us": val(r, "gpu__time_duration.sum")}
        pos += int(cnt)
        name = prefix + ("_ncu_traffic.json" if k == 0 else "_%s_ncu_traffic.json" % cfg)
        json.dump(traffic, open(name, "w"), indent=1)
    # the few metrics the roofline discussion in DESIGN.md / profiles/README.md quotes, one readable table
    KEY = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_active.avg",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
           "smsp__average_warp_latency_per_inst_issued.ratio", "local_load_requests", "smsp__inst_executed_op_local_ld.sum",
           "smsp__inst_executed_op_local_st.sum"]
    with open(prefix + "_ncu_key_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + kernels)
        for nm in KEY:
            if nm in names:
                j = names.index(nm)
                w.writerow([nm, units[j]] + [r[j] for r in rows])
    with open(prefix + "_ncu_source_mix.txt", "w") as f:
        for k in sorted(set(kernels)):
            mix = source_mix(rep, k)
            if mix is None:
                f.write("== %s: no source page\n" % k)
                continue
            ex, wf, wfi = mix
            f.write("== %s: warp-level executed instructions %d, shared-memory wavefronts %d (ideal %d)\n"
                    % (k, sum(ex.values()), sum(wf.values()), sum(wfi.values())))
            for op, c in ex.most_common(40):
                f.write("  %-22s %12d  smem wavefronts %12d (ideal %d)\n" % (op, c, wf[op], wfi[op]))
    print("wrote", prefix + "_ncu_{summary.csv,key_metrics.csv,traffic.json,source_mix.txt}")


if __name__ == "__main__":
    main()
