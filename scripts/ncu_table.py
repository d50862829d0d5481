"""Markdown table from an `ncu --set full` report:  python scripts/ncu_table.py gpurun_out/prof.ncu-rep
(reads `ncu -i <rep> --page raw --csv`; one row per profiled launch)."""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
data = [r for r in rows[2:] if len(r) == len(hdr)]  # rows[1] holds the units
col = {n: i for i, n in enumerate(hdr)}


def f(r, name, scale=1.0):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return float("nan")
    return float(r[i].replace(",", "")) * scale


units = dict(zip(hdr, rows[1]))


def to_bytes(r, name):
    v, u = f(r, name), units.get(name, "").split("/")[0]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(r, name):
    v, u = f(r, name), units.get(name, "")
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1)


print("| kernel | grid | duration us (under ncu) | dram read MB | dram write MB | DRAM % of peak | tensor pipe % | "
      "regs | dyn smem/CTA KB |")
print("|---|---|---|---|---|---|---|---|---|")
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("dfl::", "")
    grid = r[col["Grid Size"]] if "Grid Size" in col else ""
    tens = f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    if tens != tens:
        tens = f(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active")
    print(f"| {name} | {grid} | {to_us(r, 'gpu__time_duration.sum'):.1f} | {to_bytes(r, 'dram__bytes_read.sum') / 1e6:.1f} | "
          f"{to_bytes(r, 'dram__bytes_write.sum') / 1e6:.2f} | {f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{tens:.1f} | {f(r, 'launch__registers_per_thread'):.0f} | "
          f"{to_bytes(r, 'launch__shared_mem_per_block_dynamic') / 1e3:.1f} |")

# --lm-json <path>: per-launch DRAM traffic of the lm_head GEMM (+ fused argmax) for bench.py's roofline.traffic
if "--lm-json" in sys.argv:
    import json
    out = sys.argv[sys.argv.index("--lm-json") + 1]
    lm = [r for r in data if re.search(r"gemm_skinny_kernel<\d+, 1>", r[col["Kernel Name"]])]
    assert lm, "no lm_head GEMM (argmax mode) launch in this report"
    tr = [to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum") for r in lm]
    json.dump(dict(kernel="gemm_skinny_kernel<16,argmax> (lm_head)", launches=len(lm),
                   dram_bytes_per_launch=sum(tr) / len(tr),
                   dram_read_bytes=sum(to_bytes(r, "dram__bytes_read.sum") for r in lm) / len(lm),
                   dram_write_bytes=sum(to_bytes(r, "dram__bytes_write.sum") for r in lm) / len(lm),
                   duration_us_under_ncu=sum(to_us(r, "gpu__time_duration.sum") for r in lm) / len(lm),
                   source="ncu --set full --clock-control none, " + rep.split("/")[-1]), open(out, "w"), indent=1)
