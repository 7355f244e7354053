"""Turn ncu reports / launch lists into the tracked summaries under profiles/.
  python profiles/scripts/summarize_ncu.py rep OUT.md TITLE name1=path1.ncu-rep name2=path2.ncu-rep ...
      -> profiles/<round>/<name>.summary.csv (every metric of the capture) + a markdown table of the key ones
  python profiles/scripts/summarize_ncu.py launches OUT.md TITLE launches.csv
      -> per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list
Run where ncu is installed (it only READS reports here; captures are taken on the GPU box)."""
import collections, csv, os, subprocess, sys

KEYS = [
    ("duration", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("registers/thread", "launch__registers_per_thread"),
    ("dyn. shared mem / CTA", "launch__shared_mem_per_block_dynamic"),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("active threads per instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("FP64 pipe active %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("LSU data-pipe wavefronts % of peak", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
    ("shared-memory wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
    ("dram throughput % of peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("stall long_scoreboard / issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall wait / issue", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall short_scoreboard / issue", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall barrier / issue", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall not_selected / issue", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("stall math_pipe_throttle / issue", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
]


def load_rep(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    return {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def main():
    mode, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    outdir = os.path.join(os.path.dirname(os.path.abspath(out)), os.path.basename(out).split("_")[0])
    if mode == "rep":
        os.makedirs(outdir, exist_ok=True)
        names, data = [], []
        for arg in sys.argv[4:]:
            name, path = arg.split("=", 1)
            d = load_rep(path)
            names.append((name, d.get("Kernel Name", ("", ""))[1]))
            data.append(d)
            with open(os.path.join(outdir, name + ".summary.csv"), "w") as f:
                w = csv.writer(f)
                w.writerow(["metric", "unit", "value"])
                for k, (u, v) in d.items():
                    if "__" in k:
                        w.writerow([k, u, v])
        with open(out, "w") as f:
            f.write("# %s\n\nRaw metrics per kernel: `%s/*.summary.csv`.\n\n" % (title, os.path.relpath(outdir, os.path.dirname(out))))
            f.write("| metric | " + " | ".join("%s `%s`" % (n, k[:48]) for n, k in names) + " |\n")
            f.write("|---|" + "---:|" * len(names) + "\n")
            for label, key in KEYS:
                cells = []
                for d in data:
                    u, v = d.get(key, ("", "n/a"))
                    try:
                        v = "%.4g" % float(v)
                    except ValueError:
                        pass
                    cells.append("%s %s" % (v, u))
                f.write("| %s | %s |\n" % (label, " | ".join(cells)))
    else:
        rows = list(csv.reader(open(sys.argv[4])))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        hdr = rows[hi]
        kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        tot, cnt = collections.Counter(), collections.Counter()
        for r in rows[hi + 1:]:
            if len(r) <= mv:
                continue
            v = float(r[mv].replace(",", ""))
            v = v / 1e6 if r[mu] in ("ns", "nsecond") else v / 1e3 if r[mu] in ("us", "usecond") else v
            tot[r[kn]] += v
            cnt[r[kn]] += 1
        allms = sum(tot.values())
        ours = sum(v for k, v in tot.items() if "fpb::" in k)
        with open(out, "w") as f:
            f.write("# %s\n\nPer-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n" % title)
            f.write("| kernel | launches | total ms | share of all | share of this repo's kernels |\n|---|---:|---:|---:|---:|\n")
            for k, v in tot.most_common(20):
                f.write("| `%s` | %d | %.3f | %.1f%% | %s |\n" % (k[:90], cnt[k], v, 100 * v / allms,
                                                            ("%.1f%%" % (100 * v / ours)) if "fpb::" in k else "-"))


if __name__ == "__main__":
    main()
