"""Merge per-kernel figures of one `ncu --set full` capture (its `--page raw --csv` export) into
profiles/ncu_kernels.json, the file bench.py reads `roofline.traffic` from -- nothing in that file is typed by hand.

    python tools/ncu_kernels_json.py gpurun_out/r2c3_sumfac_raw.csv biquadratic:128x128x128:4lev:n1 [label]

One record per (kernel, workload): the LONGEST launch of that kernel in the capture (the finest level's).  Kernel names are shortened to the
ones the library reports (b2_asm_kernel_name) / bench.py uses."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_kernels.json")


def short_name(full):
    m = re.search(r"(\w+_kernel)\s*<([^>]*)>", full) or re.search(r"(\w+_kernel)", full)
    name = m.group(1)
    if name in ("assemble_q2_sumfac_kernel", "assemble_q2_mma_kernel") and m.lastindex == 2:
        args = [a.strip() for a in m.group(2).split(",")]
        # sumfac: <WARPS, SlotT, GAL, CSlotT>; mma: <SlotT, GAL, CSlotT>
        gal = args[2] if name == "assemble_q2_sumfac_kernel" else args[1]
        if gal in ("1", "true", "(bool)1"):
            name += "<fused Galerkin>"
    return name


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main(path, workload, label=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def get(r, key, scale_unit=True):
        if key not in ix:
            return None
        v = num(r[ix[key]])
        if v is None or not scale_unit:
            return v
        u = units[ix[key]].lower()
        mult = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}
        return v * mult.get(u, 1.0)

    db = {"kernels": []}
    if os.path.exists(OUT):
        db = json.load(open(OUT))
    seen = {}
    for r in rows[2:]:
        k = short_name(r[ix["Kernel Name"]])
        cyc = get(r, "sm__cycles_elapsed.max")
        rec = {"kernel": k, "workload": workload, "source": (label or os.path.basename(path)) + " (ncu --set full --clock-control none, one launch)",
               "full_name": r[ix["Kernel Name"]][:160],
               "time_ms": get(r, "gpu__time_duration.sum"),
               "dram_bytes_per_launch": (get(r, "dram__bytes_read.sum") or 0.0) + (get(r, "dram__bytes_write.sum") or 0.0),
               "dram_bytes_read": get(r, "dram__bytes_read.sum"), "dram_bytes_write": get(r, "dram__bytes_write.sum"),
               "fp64_pipe_pct": get(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
               "l1tex_data_pipe_pct": get(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
               "issue_active_pct": get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "dram_throughput_pct": get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
               "registers_per_thread": get(r, "launch__registers_per_thread"),
               "shared_bank_conflicts": get(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")}
        f = {op: get(r, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") for op in ("dfma", "dmul", "dadd")}
        if cyc and f["dfma"] is not None:
            rec["dfma_flop_per_launch"] = cyc * (2.0 * f["dfma"] + (f["dmul"] or 0.0) + (f["dadd"] or 0.0))
        if k in seen and seen[k]["time_ms"] >= rec["time_ms"]:
            continue
        seen[k] = rec
        db["kernels"] = [x for x in db["kernels"] if not (x["kernel"] == k and x["workload"] == workload)] + [rec]
        print(k, {a: b for a, b in rec.items() if a not in ("full_name", "source", "workload", "kernel")})
    json.dump(db, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:4])
