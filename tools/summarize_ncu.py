"""Turn the ncu outputs under gpurun_out/ into the small text summaries committed under profiles/."""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr, agg, order = None, collections.OrderedDict(), []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                k = d["Kernel Name"]
                a = agg.setdefault(k, [0, 0.0, d["Grid Size"], d["Block Size"], []])
                a[0] += 1
                a[1] += float(d["Metric Value"])
                a[4].append(float(d["Metric Value"]))
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes (the first launch of a kernel pays module load)\n")
        f.write("share%  total_us  launches  avg_us  median_us  grid  block  kernel\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            med = sorted(v[4])[len(v[4]) // 2]
            f.write(f"{100*v[1]/tot:6.1f}  {v[1]/1e3:9.1f}  {v[0]:5d}  {v[1]/v[0]/1e3:8.1f}  {med/1e3:8.1f}  {v[2]}  {v[3]}  {k[:110]}\n")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({rep}); per launch\n")
        for r in rows[2:]:
            f.write("\n== " + r[hdr.index("Kernel Name")] + "\n")
            for i, h in enumerate(hdr):
                if h in KEEP:
                    f.write(f"{h:95s} {r[i]:>16s} {units[i]}\n")


if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    (launches if kind == "launches" else full)(src, dst)
