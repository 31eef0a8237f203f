"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table for profiles/.

    python tools/launch_summary.py gpurun_out/r1_launches.csv profiles/r1_launches_bench.txt "command line"

share = fraction of the time of all libwsi_hgnn.so kernels (torch's own fill / index / cat kernels are listed, not shared).
"""
import collections
import csv
import sys

OURS = ("typed_linear", "attn_", "segment_pool", "split_bf16", "convert_operand", "row_sqnorm", "skip_mix_bwd", "adam_flat", "csr_", "scan_", "work_", "knn_", "edge_pearson",
        "rel_transform", "layernorm", "segment_combine", "skip_mix", "wgrad", "colsum", "adam", "halo", "gather_rows")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    hdr, agg = None, collections.OrderedDict()
    for r in csv.reader(open(src)):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1000 if d["Metric Unit"] == "ns" else (v * 1000 if d["Metric Unit"] == "ms" else v)
        name = d["Kernel Name"].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name[:64], [0, 0.0])
        a[0] += 1
        a[1] += v
    ours = sum(v[1] for k, v in agg.items() if any(o in k for o in OURS))
    with open(dst, "w") as f:
        f.write(f"# {dst.split('/')[-1]} - ncu --metrics gpu__time_duration.sum --clock-control none  {cmd}\n")
        f.write("# every launch of the run (warm-ups, CUDA-graph capture + replays, eager, e2e, per-kernel roofline sections);\n"
                "# times are cold-cache and serialised: compare SHARES, not absolutes.\n"
                "# share = fraction of the time of all libwsi_hgnn.so kernels.\n")
        f.write(f"{'kernel':64s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            mine = any(o in k for o in OURS)
            share = f"{v[1] / ours:7.3f}" if mine and ours else "  (torch)"
            f.write(f"{k:64s} {v[0]:8d} {v[1]:10.1f} {v[1] / v[0]:8.2f} {share}\n")
    print(open(dst).read())


if __name__ == "__main__":
    main()
