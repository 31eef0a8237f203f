"""Summarise an .ncu-rep (read with `ncu -i`, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/r1_attn_fwd.ncu-rep profiles/r1_attn_fwd.txt [traffic_key]

Prints the headline metrics of every captured launch (duration, DRAM bytes, L2/SM/tensor-pipe utilisation,
registers, occupancy) and the top stall lines of the source page; with `traffic_key` also records
dram read+write bytes per launch in profiles/traffic.json (bench.py reports it as roofline.traffic).
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.avg", "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    lines = [f"# {os.path.basename(rep)} - ncu --set full --clock-control none (per-launch, cold-cache, serialised)"]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    traffic = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"\n## launch {d.get('ID')}: {d.get('Kernel Name', '')[:100]}")
        for k in KEYS:
            if k in d:
                lines.append(f"{k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
        try:
            rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            traffic = rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]]
            lines.append(f"{'dram read+write bytes per launch':75s} {traffic:16.0f} byte")
        except (KeyError, ValueError):
            pass
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    if len(src) > 2:
        h = src[1]
        try:
            i_src, i_s, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
            stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
            data = [r for r in src[2:] if len(r) > i_s and r[i_s].isdigit()]
            tot = sum(int(r[i_s]) for r in data)
            lines.append(f"\n## top stall samples by SASS instruction (total samples {tot})")
            for r in sorted(data, key=lambda r: -int(r[i_s]))[:20]:
                st = sorted(((h[i], int(r[i])) for i in stall if r[i].isdigit() and int(r[i]) > 0), key=lambda kv: -kv[1])[:2]
                lines.append(f"{int(r[i_s]):6d} samples  executed {r[i_ex]:>9s}  {r[i_src][:80]:80s} {st}")
        except ValueError:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    if key and traffic is not None:
        p = os.path.join(os.path.dirname(out), "traffic.json")
        t = json.load(open(p)) if os.path.exists(p) else {}
        t[key] = traffic
        json.dump(t, open(p, "w"), indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
