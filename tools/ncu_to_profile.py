"""ncu captures -> profiles/r02_kernels.json (what bench.py's roofline reads).

    python tools/ncu_to_profile.py cfg5=gpurun_out/r02_cfg5_raw.csv cfg2=... cfg1=... cfg4=...

For every workload: the LAST launch of every kernel in the .ncu-rep (warm), its duration, FP64-pipe / issue / XU
utilisation, thread-level DFMA / DMUL / DADD counts, DRAM bytes; per workload the executed FP64 FLOP and DRAM bytes per
pixel*ray (tools/profile_run.py PIXEL_RAYS).  The file records the sha256 of csrc/ it was captured on; bench.py
withholds these figures when the kernels have changed since.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

PIXEL_RAYS = {"cfg5": 41 * 41 * 10 ** 6, "cfg2": 1681 * 10 ** 4, "cfg1": 2601 * 10 ** 4, "cfg4": 25 * 2601 * 10 ** 4}
M = {
    "time_us": ("gpu__time_duration.sum", 1e-3),
    "fp64_pipe_pct": ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    "xu_pipe_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
    "fma_pipe_pct": ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    "alu_pipe_pct": ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    "registers": ("launch__registers_per_thread", 1),
    "dfma": ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 1),
    "dmul": ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 1),
    "dadd": ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 1),
    "warp_inst": ("smsp__inst_executed.sum", 1),
    "dram_read": ("dram__bytes_read.sum", 1),
    "dram_write": ("dram__bytes_write.sum", 1),
}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


KERNELS = {"cfg5": ("ff_",), "cfg2": ("ff_",), "cfg1": ("nf_", "ff_pair"), "cfg4": ("nf_", "ff_pair")}
REPS = 2      # tools/ncu_all.sh runs `profile_run.py <workload> 2`


def read(rep, workload):
    """`rep`: an .ncu-rep, or the CSV of its raw page (ncu -i rep --page raw --csv).  Per kernel of the workload's
    timed map: launches per map, per-map sums of time / FP64 operation counts / DRAM bytes (mean over the captured
    maps for the counts, the last -- warm -- map for the time), utilisation figures of its longest launch."""
    if rep.endswith(".csv"):
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        if not name.startswith(KERNELS[workload]):
            continue
        k = {}
        for key, (metric, scale) in M.items():
            if metric in ix and r[ix[metric]] != "":
                v = float(r[ix[metric]].replace(",", ""))
                u = units[ix[metric]]
                if key == "time_us":
                    v = v * TO_US[u]
                elif key.startswith("dram"):
                    v = v * BYTES[u]
                k[key] = v
        k["grid"] = r[ix["Grid Size"]] if "Grid Size" in ix else ""
        k["block"] = r[ix["Block Size"]] if "Block Size" in ix else ""
        launches.setdefault(name, []).append(k)
    out = {}
    for name, ls in launches.items():
        per = max(1, len(ls) // REPS)
        last = ls[-per:]
        top = max(last, key=lambda k: k.get("time_us", 0))
        e = {key: top.get(key) for key in ("fp64_pipe_pct", "issue_active_pct", "xu_pipe_pct", "fma_pipe_pct", "alu_pipe_pct",
                                           "warps_active_pct", "registers", "grid", "block")}
        e["launches_per_map"] = per
        e["time_us"] = sum(k.get("time_us", 0) for k in last)
        for key in ("dfma", "dmul", "dadd", "warp_inst", "dram_read", "dram_write"):
            e[key] = sum(k.get(key, 0) for k in last)
        out[name] = e
    return out


def main():
    prof = {"csrc_sha256": bench.csrc_hash(), "how": "tools/ncu_all.sh: ncu --set full + thread-level DFMA/DMUL/DADD counters, --clock-control none, on `tools/profile_run.py <workload> 2`; figures of the second (warm) map",
            "workloads": {}}
    if os.path.exists(bench.PROFILE_JSON):
        old = json.load(open(bench.PROFILE_JSON))
        if old.get("csrc_sha256") == prof["csrc_sha256"]:
            prof["workloads"] = old.get("workloads", {})
    for arg in sys.argv[1:]:
        wl, rep = arg.split("=", 1)
        kernels = read(rep, wl)
        flop = sum(2 * k.get("dfma", 0) + k.get("dmul", 0) + k.get("dadd", 0) for k in kernels.values())
        dram = sum(k.get("dram_read", 0) + k.get("dram_write", 0) for k in kernels.values())
        prof["workloads"][wl] = {"pixel_rays_captured": PIXEL_RAYS[wl], "fp64_flop_per_pixel_ray": flop / PIXEL_RAYS[wl],
                                 "dram_bytes_per_pixel_ray": dram / PIXEL_RAYS[wl], "captured_on": os.path.basename(rep),
                                 "kernels": kernels}
        print(wl, "executed FP64 FLOP per pixel*ray %.2f, DRAM B per pixel*ray %.4g" % (flop / PIXEL_RAYS[wl], dram / PIXEL_RAYS[wl]))
        for n, k in kernels.items():
            print("   %-28s %9.1f us  fp64 pipe %5.1f %%  issue %5.1f %%  regs %3d" % (
                n[:28], k.get("time_us", 0), k.get("fp64_pipe_pct", 0), k.get("issue_active_pct", 0), int(k.get("registers", 0))))
    json.dump(prof, open(bench.PROFILE_JSON, "w"), indent=1)


if __name__ == "__main__":
    main()
