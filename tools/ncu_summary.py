"""selected metrics of an ncu report as JSON (run where ncu is installed; reads with `ncu -i ... --page raw --csv`)"""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__inst_executed_pipe_fp64.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "?")
        res[name] = {k: {"value": d[k], "unit": units[hdr.index(k)]} for k in KEYS if k in d}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1)[:3000])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
