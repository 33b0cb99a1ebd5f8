"""Summarise an `ncu --page raw --csv` dump (read here, on the CPU box) into the small JSON files kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv --out profiles/r01_ncu_full_c2_bf16.json \
        [--traffic profiles/igemm_traffic_c2_bf16.json --kernel igemm_tc]

Per launch: kernel name, grid/block, duration, DRAM bytes read/written, tensor-pipe and SM throughput percentages,
registers. `--traffic` additionally writes the mean DRAM traffic per launch of the kernels matching `--kernel`
(bench.py's roofline.traffic reads that file). Test/profiling infrastructure, not part of the product.
"""
import argparse
import csv
import json
import sys

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

KEEP = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__cycles_active.avg": "smsp_cycles_active_avg",
}


def parse(path):
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if r]
    # skip ncu's ==PROF== banner lines
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    header, units, data = rows[start], rows[start + 1], rows[start + 2:]
    col = {n: i for i, n in enumerate(header)}
    out = []
    for r in data:
        if len(r) < len(header):
            continue
        rec = {"id": int(r[col["ID"]]), "kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]],
               "block": r[col["Block Size"]]}
        for name, key in KEEP.items():
            if name not in col or r[col[name]] == "":
                continue
            try:
                v = float(r[col[name]].replace(",", ""))
            except ValueError:
                continue
            u = units[col[name]]
            if key in ("dram_read_bytes", "dram_write_bytes", "duration_us"):
                v *= UNIT_SCALE.get(u, 1.0)
            rec[key] = v
        out.append(rec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic")
    ap.add_argument("--kernel", default="igemm_tc")
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    recs = parse(args.raw_csv)
    summary = {"source": args.raw_csv, "note": args.note, "launches": recs}
    sel = [r for r in recs if any(k in r["kernel"] for k in args.kernel.split(",")) and "dram_read_bytes" in r]
    if sel:
        tot_b = sum(r["dram_read_bytes"] + r["dram_write_bytes"] for r in sel)
        tot_us = sum(r.get("duration_us", 0.0) for r in sel)
        summary["selected"] = {
            "kernel": args.kernel, "launches": len(sel), "dram_bytes_total": tot_b,
            "dram_bytes_per_launch": tot_b / len(sel), "duration_us_total": tot_us,
            "tensor_pipe_pct_active_time_weighted":
                sum(r.get("tensor_pipe_pct_active", 0.0) * r.get("duration_us", 0.0) for r in sel) / max(tot_us, 1e-9),
        }
        if args.traffic:
            with open(args.traffic, "w") as f:
                json.dump({"kernel": args.kernel, "launches": len(sel), "dram_bytes_per_launch": tot_b / len(sel),
                           "dram_bytes_total_per_step": tot_b, "from": args.out, "note": args.note}, f, indent=1)
    with open(args.out, "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary.get("selected", {}), indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
