"""Turn the ncu outputs of scripts/final_r2.sh (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py r1
"""
import csv, json, os, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def launch_list():
    rows = []
    with open(os.path.join(G, "launches.csv")) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    for r in rd:
        if r[im] == "gpu__time_duration.sum":
            rows.append((r[ik], float(r[iv].replace(",", ""))))
    agg = collections.OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(G, "launches.csv")) as f, open(os.path.join(P, f"{tag}_launches.csv"), "w") as o:
        o.write(f.read())
    out = [f"# Round {tag[1:]} -- launch list of `bench.py` (ncu, per-launch device time)", "",
           "Command (run on a B200 through gpurun, directly after the same command had exited 0 without ncu; scripts/final_r2.sh):", "",
           "    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \\",
           "        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-fit --no-secondary", "",
           f"Raw list: `profiles/{tag}_launches.csv`.  Times under ncu are serialised and cold-cache; compare shares.", "",
           "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:110]}` | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.2f} % |")
    smo = sum(a[1] for k, a in agg.items() if "smo_persistent" in k)
    out += ["", f"`b2svm::smo_persistent` is {100 * smo / tot:.2f} % of the device time of the run: one launch = the whole",
            "`for n_iter in range(max_iter)` loop of the reference (select + two Q columns + step + gradient update per iteration);",
            "`<float, 32, 0, 0>` = float32 rows of 32 16-byte chunks, column cache compiled out, plain C-SVC shape.  Both legs of",
            "the bench launch it: the resident-input leg directly, the end-to-end leg through `BiKernelSVC(max_iter=20000).fit`",
            "(the estimator's automatic cache policy keeps the cache off when max_iter < n: no column could be reused).",
            "Everything else (flag refresh, state init, row norms, bias statistics, the torch fills of the L2 flush) is noise,",
            "in agreement with bench.py's kernel_ms_per_launch vs ms_per_step."]
    open(os.path.join(P, f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")
    return agg


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def pick(m, u, names):
    out = []
    for n in names:
        if n in m:
            out.append(f"| {n} | {m[n]} {u[n]} |")
    return out


def full_capture(rep, title, md, extra_lines, names):
    m, u = raw_metrics(rep)
    out = [title, ""] + extra_lines + ["", "| metric | value |", "|---|---|"] + pick(m, u, names)
    stalls = {k.split("smsp__average_warp")[-1]: v for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")}
    top = sorted(((float(v), k.replace("s_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k, v in stalls.items() if v), reverse=True)[:7]
    out += ["", "Warp stall reasons (per issue-active): " + ", ".join(f"{k} {v:.2f}" for v, k in top) + "."]
    open(os.path.join(P, md), "w").write("\n".join(out) + "\n")
    return m


if __name__ == "__main__":
    launch_list()
    names = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
             "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
             "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
             "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
             "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
             "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    m = full_capture(os.path.join(G, f"prof_{tag}_smo.ncu-rep"),
                     f"# Round {tag[1:]} -- `ncu --set full` of the persistent SMO kernel (C3: n=200 000, d=128, float32, RBF, cache off)", f"{tag}_smo_persistent_ncu_full.md",
                     ["Command (gpurun, one B200; the plain run of the same command exited 0 directly before; scripts/final_r2.sh):", "",
                      "    ncu --set full --clock-control none --import-source on -k regex:smo_persistent -s 1 -c 1 \\",
                      f"        -o gpurun_out/prof_{tag}_smo python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-fit --no-secondary", "",
                      "One captured launch = 20 000 SMO iterations (148 CTAs x 384 threads: 11 consumer warps + 1 TMA producer warp)."], names)
    rd, wr = float(m["dram__bytes_read.sum"]), float(m["dram__bytes_write.sum"])
    unit = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "byte": 1.0}
    _, u = raw_metrics(os.path.join(G, f"prof_{tag}_smo.ncu-rep"))
    rd *= unit[u["dram__bytes_read.sum"]]
    wr *= unit[u["dram__bytes_write.sum"]]
    json.dump({"source": f"profiles/{tag}_smo_persistent_ncu_full.md (ncu --set full, one launch of 20000 iterations, n=200000 d=128)",
               "iterations_in_capture": 20000, "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_iteration": (rd + wr) / 20000}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    if os.path.exists(os.path.join(G, f"prof_{tag}_matvec.ncu-rep")):
        full_capture(os.path.join(G, f"prof_{tag}_matvec.ncu-rep"),
                     f"# Round {tag[1:]} -- `ncu --set full` of the tensor-core kernel mat-vec (K7: 400 000 queries x 200 000 support rows, d=32, RBF)", f"{tag}_matvec_tc_ncu_full.md",
                     ["Command (gpurun, one B200, after the plain run exited 0; scripts/final_r2.sh):", "",
                      "    ncu --set full --clock-control none --import-source on -k regex:matvec_tc_kernel -s 1 -c 1 \\",
                      f"        -o gpurun_out/prof_{tag}_matvec python scripts/ncu_matvec_target.py", "",
                      "One launch = 8e10 kernel evaluations: 3125 query tiles x 2 support slices, 576 threads (TMA warp, MMA warp, 16 epilogue warps)."], names)
    print("profiles written")
