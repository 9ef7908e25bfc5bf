"""Key metrics per kernel from ncu reports (run where ncu is installed; no GPU needed):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt
    python tools/ncu_summary.py --roofline-json profiles/rNN_gemm_roofline.json --kernel gemm_tc_kernel \
           --batch 128 rep1.ncu-rep [rep2.ncu-rep ...] [summary.txt ...]

The second form writes the per-launch DRAM traffic and the time-weighted tensor-pipe activity of every captured launch of
`--kernel` into one JSON file, which bench.py reads for `roofline.traffic` / `tensor_pipe_active_pct` (so those fields
come from a named capture, not from literals in the source).  Inputs may be .ncu-rep files or text summaries written
by the first form (their numbers are parsed back)."""
import csv
import json
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_lsu.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]

_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
          "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def _num(v):
    return float(v.replace(",", ""))


def rows_of(path):
    """[(kernel name, {metric: (value, unit)})] from a .ncu-rep or a text summary of this tool."""
    out = []
    if path.endswith(".ncu-rep") or path.endswith(".csv"):   # a report, or its `--page raw --csv` dump made on the GPU box
        if path.endswith(".csv"):
            txt = open(path).read()
        else:
            txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            m = {}
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    m[w] = (r[i], units[i])
            out.append((r[hdr.index("Kernel Name")], m))
        return out
    cur = None
    with open(path) as f:
        for line in f:
            if line.startswith("## "):
                cur = (line[3:].strip(), {})
                out.append(cur)
            elif cur is not None and line.startswith("   "):
                parts = line.split()
                if len(parts) >= 2:
                    cur[1][parts[0]] = (parts[1], parts[2] if len(parts) > 2 else "")
    return out


def text_summary(paths):
    for path in paths:
        print(f"# {path}")
        for name, m in rows_of(path):
            print(f"## {name[:110]}")
            for w in WANT:
                if w in m:
                    print(f"   {w:95s} {m[w][0]:>16s} {m[w][1]}")


def roofline_json(paths, kernel, batch, out_path):
    launches = []
    for path in paths:
        for name, m in rows_of(path):
            if kernel not in name:
                continue
            t = _num(m["gpu__time_duration.sum"][0]) * _SCALE.get(m["gpu__time_duration.sum"][1], 1.0)
            rd = _num(m["dram__bytes_read.sum"][0]) * _SCALE.get(m["dram__bytes_read.sum"][1], 1.0)
            wr = _num(m["dram__bytes_write.sum"][0]) * _SCALE.get(m["dram__bytes_write.sum"][1], 1.0)
            tp = _num(m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0])
            launches.append({"kernel": re.sub(r"\(.*", "", name), "time_us": t, "dram_bytes": rd + wr, "tensor_pipe_active_pct": tp,
                             "source": path})
    if not launches:
        raise SystemExit(f"no launch of {kernel} in {paths}")
    tsum = sum(l["time_us"] for l in launches)
    doc = {"kernel": kernel, "batch": batch, "captures": paths, "n_launches": len(launches),
           "traffic_bytes_per_launch": sum(l["dram_bytes"] for l in launches) / len(launches),
           "tensor_pipe_active_pct_time_weighted": sum(l["tensor_pipe_active_pct"] * l["time_us"] for l in launches) / tsum,
           "time_us_per_launch_under_ncu": tsum / len(launches),
           "note": "ncu --set full --clock-control none; per-launch times are cold-cache and serialised",
           "launches": launches}
    with open(out_path, "w") as f:
        json.dump(doc, f, indent=1)
    print(f"wrote {out_path}: {len(launches)} launches, traffic {doc['traffic_bytes_per_launch'] / 1e6:.1f} MB/launch, "
          f"tensor pipe {doc['tensor_pipe_active_pct_time_weighted']:.1f} %")


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--roofline-json":
        out_path, kernel, batch, rest = args[1], "gemm_tc_kernel", 128, args[2:]
        while rest and rest[0].startswith("--"):
            if rest[0] == "--kernel":
                kernel = rest[1]
            elif rest[0] == "--batch":
                batch = int(rest[1])
            rest = rest[2:]
        roofline_json(rest, kernel, batch, out_path)
    else:
        text_summary(args)
