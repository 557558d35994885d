#!/usr/bin/env python
"""Per-kernel pipe counters of ncu --set full reports -> profiles/traffic.json (read by bench.py for the per-stage fractions).

    tools/ncu_pipes.py <images per launch> <out.json> <report.ncu-rep> [<report.ncu-rep> ...]

For every extractor / stereo stage (one launch per distinct kernel and grid; the pyramid and stereo stages are sums over their
kernels) it records, per launch: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), warp instructions
(smsp__inst_executed.sum), ALU- and FMA-pipe warp instructions (sm__inst_executed_pipe_{alu,fma} percent of peak x 2 warp
instructions per active SM cycle -- the ALU pipe takes one warp instruction every second cycle per scheduler, four schedulers),
shared-memory wavefronts (l1tex__data_pipe_lsu_wavefronts_mem_shared.sum), ncu's own percent-of-peak figures and the binding pipe
(the largest of them)."""
import csv, io, json, subprocess, sys

STAGE = {"k_fast_cells": "fast", "k_quadtree": "quadtree", "k_blur_tile": "blur", "k_describe": "describe",
         "k_resize_tile": "pyramid", "k_resize_chain": "pyramid", "k_resize": "pyramid",
         "k_stereo_rows": "stereo", "k_stereo_match": "stereo", "k_stereo_filter": "stereo"}


def main():
    per_img, out_path, reps = int(sys.argv[1]), sys.argv[2], sys.argv[3:]
    seen, stages = set(), {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr = rows[0]
        ix = {h: i for i, h in enumerate(hdr)}

        def val(r, k):
            return float(r[ix[k]].replace(",", "")) if k in ix and r[ix[k]] not in ("", "n/a") else 0.0
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
            if name not in STAGE:
                continue
            key = (name, r[ix["launch__grid_size"]])
            if key in seen:                      # one launch per distinct kernel and grid
                continue
            seen.add(key)
            cyc = val(r, "sm__cycles_active.sum")
            unit = rows[1][ix["gpu__time_duration.sum"]]
            us = val(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
            k = {"kernel": name, "grid": int(float(key[1])), "ncu_us": us,
                 "dram_bytes": val(r, "dram__bytes_read.sum") * _scale(rows[1][ix["dram__bytes_read.sum"]]) +
                               val(r, "dram__bytes_write.sum") * _scale(rows[1][ix["dram__bytes_write.sum"]]),
                 "warp_instructions": val(r, "smsp__inst_executed.sum"),
                 "alu_instructions": val(r, "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active") / 100 * 2 * cyc,
                 "fma_instructions": val(r, "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active") / 100 * 2 * cyc,
                 "lsu_wavefronts": val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                 "sm_cycles_active": cyc,
                 "pct": {"issue": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                         "alu": val(r, "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
                         "fma": val(r, "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active"),
                         "lsu_wavefronts": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                         "dram": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}}
            stages.setdefault(STAGE[name], []).append(k)
    out = {"_images_per_launch": per_img, "_source": " ".join(reps) + ": ncu --set full, one launch per distinct kernel and grid",
           "_pipes": {}}
    for st, ks in stages.items():
        tot = lambda f: sum(k[f] for k in ks)
        out[st] = tot("dram_bytes")
        us = tot("ncu_us")
        pct = {p: sum(k["pct"][p] * k["ncu_us"] for k in ks) / us for p in ks[0]["pct"]}      # time-weighted over the stage's kernels
        binding = max(("issue", "alu", "fma", "lsu_wavefronts", "dram"), key=lambda p: pct[p])
        out["_pipes"][st] = {"kernels": [f'{k["kernel"]}[{k["grid"]}]' for k in ks], "ncu_us": us,
                             "warp_instructions": tot("warp_instructions"), "alu_instructions": tot("alu_instructions"),
                             "fma_instructions": tot("fma_instructions"), "lsu_wavefronts": tot("lsu_wavefronts"),
                             "ncu_pct_of_peak": {p: round(v, 1) for p, v in pct.items()}, "binding": binding}
    json.dump(out, open(out_path, "w"), indent=1)
    for st, p in out["_pipes"].items():
        print(f'{st:9s} {p["ncu_us"]:7.1f} us  inst {p["warp_instructions"] / 1e6:6.1f} M  alu {p["alu_instructions"] / 1e6:6.1f} M  '
              f'fma {p["fma_instructions"] / 1e6:6.1f} M  wavefronts {p["lsu_wavefronts"] / 1e6:6.1f} M  dram {out[st] / 1e6:6.1f} MB  '
              f'{p["ncu_pct_of_peak"]}  binding: {p["binding"]}')


def _scale(unit):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


if __name__ == "__main__":
    main()
