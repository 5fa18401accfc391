"""Turns ncu output brought back from gpurun into the summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/X_launches.csv profiles/X_launch_summary.csv [steps]
        per-kernel launch count / total time / share over the LAST `steps` (default 2) bench steps of a
        `ncu --metrics gpu__time_duration.sum --clock-control none --csv` log (a step ends with the
        merge_partials kernel of the search); also copies the raw launches of those steps next to it.
    python tools/ncu_summary.py full profiles/X_ncu_summary.md TITLE=rep.ncu-rep [TITLE=rep ...]
        selected metrics of every kernel in each `ncu --set full` report (read with `ncu -i ... --page raw --csv`).
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ["launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "launch__registers_per_thread",
           "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.avg.per_second"]


def short(name: str) -> str:
    name = name.replace("<unnamed>::", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    name = re.sub(r"<.*", "", name)
    return name.split("::")[-1] if "CUB" not in name and "cub" not in name else name


def read_launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        out.append((row["Kernel Name"], v))
    return out


def launches(src, dst, steps=2):
    rows = read_launches(src)
    # a bench step runs from the encoder's embed_kernel to the LAST merge_partials kernel before the next
    # embed_kernel (the two-stage scan merges three times per step); spans without a scan kernel (the
    # extra encode the bench does after its warm-up) are not steps
    begins = [i for i, (n, _) in enumerate(rows) if "embed_kernel" in n] + [len(rows)]
    spans = []
    for b, nb in zip(begins[:-1], begins[1:]):
        merges = [i for i in range(b, nb) if "merge_partials" in rows[i][0]]
        if merges and any("ivf_scan" in rows[i][0] for i in range(b, nb)):
            spans.append((b, merges[-1] + 1))
    assert len(spans) >= steps, f"need {steps} search steps in the log, found {len(spans)}"
    sel = []
    for b, e in spans[-steps:]:
        sel += rows[b:e]
    agg = collections.OrderedDict()
    for n, v in sel:
        a = agg.setdefault(short(n), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"kernel,launches({steps} steps),total_us,share\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{c},{t:.1f},{t / tot:.4f}\n")
        f.write(f"TOTAL,{len(sel)},{tot:.1f},1.0\n")
    with open(dst.replace("_launch_summary", "_launches") if "_launch_summary" in dst else dst + ".raw", "w") as f:
        f.write("kernel,us\n")
        for n, v in sel:
            f.write(f"\"{n}\",{v:.3f}\n")
    print(open(dst).read())


def full(dst, pairs):
    with open(dst, "w") as f:
        f.write("# ncu --set full summaries (B200)\n")
        for title, rep in pairs:
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
            rd = list(csv.reader(io.StringIO(raw)))
            hdr, units = rd[0], rd[1]
            f.write(f"\n## {title}\nsource: ncu --set full --clock-control none --import-source on ({rep.split('/')[-1]})\n")
            for row in rd[2:]:
                rec = dict(zip(hdr, row))
                f.write(f"\nKernel Name: {rec.get('Kernel Name')}\n")
                for m in METRICS:
                    if m in rec:
                        f.write(f"{m}: {rec[m]} {units[hdr.index(m)]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 2)
    else:
        full(sys.argv[2], [a.rsplit("=", 1) for a in sys.argv[3:]])
