"""Summarise an ncu report (read here, no GPU needed) into a small JSON for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.json
"""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "lts__t_bytes.sum", "smsp__inst_executed.sum"]

def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    launches = []
    for r in rows[2:]:
        entry = {"kernel": r[name_col]}
        for i, h in enumerate(hdr):
            for k in KEYS:
                if h == k or h.endswith("." + k):
                    try:
                        entry[k + " [" + units[i] + "]"] = float(r[i])
                    except ValueError:
                        entry[k] = r[i]
        launches.append(entry)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    stalls = {}
    if len(srows) > 2:
        sh = srows[1]
        cols = [i for i, h in enumerate(sh) if h.startswith("stall_") and "Not Issued" not in h]
        seen = set()
        for r in srows[2:]:
            if len(r) < len(sh) or r[0] in seen:
                continue
            seen.add(r[0])
            for i in cols:
                try:
                    stalls[sh[i]] = stalls.get(sh[i], 0) + int(r[i])
                except ValueError:
                    pass
        tot = sum(stalls.values()) or 1
        stalls = {k: round(100.0 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda x: -x[1]) if v}
    json.dump({"report": rep, "launches": launches, "warp_stall_sampling_pct_first_kernel": stalls}, open(out, "w"), indent=1)
    print(json.dumps(launches[0], indent=1)); print(stalls)

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
