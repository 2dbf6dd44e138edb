"""Turn .ncu-rep files into a compact markdown table of the metrics the profiling recipe names."""
import csv, subprocess, sys, io, json, os

PEAKS = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic"]


def rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    for r in rd[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        yield d


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return float("nan")


def main(paths):
    print("| kernel | grid | regs | time us | tensor pipe % | SM % | DRAM % | L2 % | DRAM MB (r+w) | achieved GB/s |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for p in paths:
        for d in rows(p):
            name = d["Kernel Name"][0].split("(")[0][-60:]
            t = num(d["gpu__time_duration.sum"][0])
            tu = d["gpu__time_duration.sum"][1]
            us = t * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(tu, 1)
            def b(k):
                v, u = d[k]
                return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            mb = (b("dram__bytes_read.sum") + b("dram__bytes_write.sum")) / 1e6
            print(f"| {name} | {d['launch__grid_size'][0]} | {d['launch__registers_per_thread'][0]} | {us:.1f} | "
                  f"{num(d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'][0]):.1f} | "
                  f"{num(d['sm__throughput.avg.pct_of_peak_sustained_elapsed'][0]):.1f} | "
                  f"{num(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0]):.1f} | "
                  f"{num(d['lts__throughput.avg.pct_of_peak_sustained_elapsed'][0]):.1f} | {mb:.1f} | {mb / us * 1e3:.0f} |")
    print(f"\nMeasured peaks used for fractions elsewhere: HBM copy {PEAKS['hbm_gbs']} GB/s, cuBLAS bf16 {PEAKS['bf16_tflops']} TFLOP/s (MEASURED_PEAKS.json).")


if __name__ == "__main__":
    main(sys.argv[1:])
