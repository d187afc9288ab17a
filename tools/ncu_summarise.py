#!/usr/bin/env python
"""Turn the `ncu --set full` captures of the sweep kernel (one .ncu-rep per workload) into the files kept under
profiles/: the full metric list per workload (<prefix>_<workload>_ncu.csv: metric,unit,launch0), a short summary
(<prefix>_ncu_summary.txt), the per-launch DRAM traffic bench.py reports as roofline.traffic
(profiles/r02_ncu_dram_traffic.json) and the instructions with the most stall samples (<prefix>_<workload>_source_top.txt).

    python tools/ncu_summarise.py gpurun_out/r02f/prof_final_ profiles/r02_final

Runs here (no GPU): `ncu -i` only reads the report."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_config_size",
    "launch__shared_mem_per_block_allocated", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    src_prefix, dst_prefix = sys.argv[1], sys.argv[2]
    traffic = {}
    summary = ["ncu --set full --clock-control none, ONE launch of cvr_spmv_tile_kernel per workload (cold, serialised), final round-2 code;",
               "command: tools/gpu_batch_r02_final.sh (tools/kernel_ab.py --variants auto under ncu).  Full metric lists: "
               + os.path.basename(dst_prefix) + "_<workload>_ncu.csv", ""]
    for w in ("rmat24", "web", "road", "fem"):
        rep = f"{src_prefix}{w}.ncu-rep"
        if not os.path.exists(rep):
            continue
        rows = ncu_csv(rep, "raw")
        names, units, vals = rows[0], rows[1], rows[2]
        with open(f"{dst_prefix}_{w}_ncu.csv", "w") as f:
            f.write("metric,unit,launch0\n")
            for n, u, v in zip(names, units, vals):
                f.write(f"{n},{u},{v}\n")
        m = {n: (u, v) for n, u, v in zip(names, units, vals)}
        kernel = m.get("Kernel Name", ("", "?"))[1]
        summary.append(f"== {w}   ({kernel[:90]})")
        for k in KEYS:
            if k in m:
                summary.append(f"  {k:75s} {m[k][1]} {m[k][0]}")
        stalls = sorted(((float(v[1].replace(',', '')), k.split("issue_stalled_")[1].split("_per_warp")[0]) for k, v in m.items()
                         if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                         and "not_issued" not in k and v[1] not in ("", "n/a")), reverse=True)[:6]
        if stalls:
            summary.append("  top stall reasons (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls))
        rd = float(m["dram__bytes_read.sum"][1].replace(',', '')) * UNIT_SCALE[m["dram__bytes_read.sum"][0]]
        wr = float(m["dram__bytes_write.sum"][1].replace(',', '')) * UNIT_SCALE[m["dram__bytes_write.sum"][0]]
        traffic[w] = {"bytes": rd + wr, "source": f"profiles/{os.path.basename(dst_prefix)}_{w}_ncu.csv (ncu --set full, one sweep launch, {kernel[:70]}...)"}
        summary.append("")
        # source page: instructions with the most stall samples
        rows = ncu_csv(rep, "source")
        hdr, data = rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        top = sorted(enumerate(data), key=lambda t: -int(t[1][ix["# Samples"]]))[:25]
        with open(f"{dst_prefix}_{w}_source_top.txt", "w") as f:
            f.write(f"ncu --set full --import-source on, {kernel[:100]} on {w}: {tot} stall samples, {len(data)} SASS instructions\n")
            f.write("index  share  executions  SASS  [top stall reasons]\n")
            for i, r in top:
                s = int(r[ix["# Samples"]])
                reasons = sorted(((int(r[ix[h]]), h[6:]) for h in st), reverse=True)[:2]
                f.write(f"{i:5d} {100 * s / tot:5.1f}% exec={r[ix['Instructions Executed']]:>9} {r[ix['Source']].strip()[:72]:72s} {reasons}\n")
    with open(f"{dst_prefix}_ncu_summary.txt", "w") as f:
        f.write("\n".join(summary) + "\n")
    with open(os.path.join(ROOT, "profiles", "r02_ncu_dram_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print("\n".join(summary))


if __name__ == "__main__":
    main()
