"""Summarise an ncu report (`ncu --set full -o X`) into a small JSON: per kernel the duration, DRAM / L2 traffic, pipe
utilisation and occupancy figures quoted in DESIGN.md.  `python tools/ncu_summary.py X.ncu-rep > summary.json`"""
import csv
import io
import json
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_sectors_mem_global_op_tma_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        d = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = (r[i] + " " + units[i]).strip()
        key, n = name, 2
        while key in res:
            key = "%s #%d" % (name, n)
            n += 1
        res[key] = d
    print(json.dumps(res, indent=1))
    return res


def traffic(res, steps):
    """profiles/rNN/traffic.json for bench.py: DRAM bytes per launch of each kernel (mean over its launches in the
    capture) and per step, tagged with the hash of the CUDA sources the capture was taken from."""
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    per, cnt = {}, {}
    for key, d in res.items():
        name = key.split(" #")[0]
        short = name.split("(")[0].replace("rb::", "").replace("void ", "").replace("<unnamed>::", "").strip()
        # the names bench.py's CUDA-event attribution uses: the forward kernel's two template instances are told apart,
        # every other kernel drops its template arguments
        if short.startswith("joint_tc"):
            short = short.replace("_kernel<2>", "_kernel<fwd+keep>").replace("_kernel<0>", "_kernel<fwd>")
        else:
            short = re.sub(r"<[^>]*>$", "", short)
        def num(k):
            v = d.get(k, "0").split()
            x = float(v[0].replace(",", "")) if v else 0.0
            u = v[1].lower() if len(v) > 1 else "byte"
            return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        b = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        per[short] = per.get(short, 0.0) + b
        cnt[short] = cnt.get(short, 0) + 1
    return {"csrc_sha16": bench.csrc_sha16(), "kernels": {k: per[k] / cnt[k] for k in per},
            "step": sum(per.values()) / max(steps, 1), "launches_in_capture": cnt, "steps_in_capture": steps}


if __name__ == "__main__":
    # `--from-json summary.json <steps> traffic.json`: rebuild traffic.json from a summary written earlier (the report itself
    # is too large to travel back from the GPU box)
    if sys.argv[1] == "--from-json":
        json.dump(traffic(json.load(open(sys.argv[2])), int(sys.argv[3])), open(sys.argv[4], "w"), indent=1)
        sys.exit(0)
    r = main(sys.argv[1])
    if len(sys.argv) >= 4:      # ncu_summary.py X.ncu-rep <steps captured> traffic.json
        json.dump(traffic(r, int(sys.argv[2])), open(sys.argv[3], "w"), indent=1)
