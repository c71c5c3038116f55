"""Turn the ncu reports captured by tools/capture_profiles.sh (gpurun_out/prof_<name>_<tag>.ncu-rep) into the tracked
text summaries under profiles/ and profiles/traffic.json (DRAM bytes per unit of work, read by bench.py)."""
import collections, csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def raw_metrics(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units, val = rows[0], rows[1], rows[2]
    return {h: (val[i], units[i]) for i, h in enumerate(hdr)}


def stalls(rep, top=14):
    rows = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot, lines = collections.Counter(), []
    for r in rows[2:]:
        if len(r) < len(hdr) or not (r[idx["# Samples"]] or "0").isdigit():
            continue
        for h in cols:
            tot[h] += int(r[idx[h]] or 0)
        d = {h[6:]: int(r[idx[h]] or 0) for h in cols if int(r[idx[h]] or 0) > 0}
        lines.append((int(r[idx["# Samples"]] or 0), r[idx["Source"]].strip(), d, int(r[idx["Instructions Executed"]] or 0)))
    n = max(sum(tot.values()), 1)
    out = ["warp-state samples: %d" % n]
    out += ["  %-26s %8d  %.3f" % (h, v, v / n) for h, v in tot.most_common(9)]
    out.append("hottest SASS lines (samples, executed, instruction, top stall reasons):")
    for i, (s_, src, d, ie) in sorted(enumerate(lines), key=lambda x: -x[1][0])[:top]:
        out.append("  #%5d %7d %10d  %-62s %s" % (i, s_, ie, src[:62], sorted(d.items(), key=lambda x: -x[1])[:2]))
    return out


def main():
    traffic = {}
    units = {"fused": ("k_reeval_fused_dram_bytes_per_walker_refresh", 1024),
             "inverse": ("k_inverse_v4_dram_bytes_per_matrix", 1024), "gemm": ("k_gemm_W_dmma_dram_bytes_per_matrix", 2048),
             "flush": ("k_flush_wb_dram_bytes_per_walker_flush", None),
             # 972 sites (tools/capture_profiles_r2.sh B: prof_refresh.py 18 512 -> 512 matrices per inverse launch, 1024 per product launch)
             "inverse972": ("k_inverse_v4_dram_bytes_per_matrix_972", 512), "gemm972": ("k_gemm_W_dmma_dram_bytes_per_matrix_972", 1024),
             # session 3 (tools/capture_profiles_r3.sh): the initial refresh of quick_bench's 4096 ComplexF64 walkers (both species
             # in one launch); prof_refresh.py 18 512 -> 1024 matrices per cluster-inverse launch
             "invclc": ("k_inverse_cl_c_dram_bytes_per_walker_refresh", 4096), "gemmc": ("k_gemm_W_dmma_c_dram_bytes_per_walker_refresh", 4096),
             "invcl972": ("k_inverse_cl_dram_bytes_per_matrix_972", 1024)}
    for name in ("fused", "inverse", "gemm", "flush", "decide", "measure", "gather", "resident", "inverse972", "gemm972", "flush972", "flushc",
                 "invclc", "gemmc", "flushdc", "invcl972", "fused192"):
        rep = os.path.join(ROOT, "gpurun_out", "prof_%s_%s.ncu-rep" % (name, TAG))
        if not os.path.exists(rep):
            print("missing", rep)
            continue
        m = raw_metrics(rep)
        out = ["# ncu --set full --clock-control none --import-source on (one launch; see tools/capture_profiles.sh), round tag %s" % TAG,
               "# kernel: %s" % m.get("Kernel Name", ("?", ""))[0]]
        for k in KEYS:
            if k in m:
                out.append("%-70s %16s %s" % (k, m[k][0], m[k][1]))
        rd = float(m["dram__bytes_read.sum"][0].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_read.sum"][1]]
        wr = float(m["dram__bytes_write.sum"][0].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_write.sum"][1]]
        out.append("dram bytes (read + write) of this launch: %.0f" % (rd + wr))
        if name in units:
            key, n_units = units[name]
            if n_units:
                traffic[key] = (rd + wr) / n_units
                out.append("%s = %.0f  (%d units in this launch)" % (key, traffic[key], n_units))
            else:
                # walkers flushed in the captured launch, from the executed count of a main-loop DMMA line:
                # items x 9 warps x 9 loop trips (27 column tiles, 3 per trip), 4 items (2 species x 2 row blocks) per walker
                rows = list(csv.reader(io.StringIO(ncu(rep, "source"))))
                hdr = rows[1]
                ix = {h: i for i, h in enumerate(hdr)}
                ex = [int(r[ix["Instructions Executed"]] or 0) for r in rows[2:] if len(r) >= len(hdr) and "DMMA" in r[ix["Source"]]]
                walkers = collections.Counter(ex).most_common(1)[0][0] / 81.0 / 4.0
                traffic[key] = (rd + wr) / walkers
                out.append("%s = %.0f  (%.0f walkers flushed in this launch; algorithmic 16 ns^2 = 2985984; dirty lines still in L2 at kernel end are not counted)" % (key, traffic[key], walkers))
        out += stalls(rep)
        with open(os.path.join(ROOT, "profiles", "%s_%s_ncu_summary.txt" % (TAG, name)), "w") as f:
            f.write("\n".join(out) + "\n")
        print("\n".join(out[:40]))
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    old = json.load(open(tj)) if os.path.exists(tj) else {}
    old.update(traffic)
    json.dump(old, open(tj, "w"), indent=1)


if __name__ == "__main__":
    main()
