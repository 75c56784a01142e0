"""Turn ncu artefacts from gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.txt>   # per-kernel count / total / share
  python tools/ncu_summary.py report   <file.ncu-rep> <out.txt>   # key metrics of each profiled launch
  python tools/ncu_summary.py traffic  <file.ncu-rep> <out.json> [algorithmic bytes per launch | "xCTA:<bytes per CTA>"] [note]
                                                                   # DRAM bytes per launch (bench.py reads this)
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name[-110:]


def launches(path, out):
    rows = [l for l in open(path, newline="") if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot, cnt = defaultdict(float), defaultdict(int)
    unit = "ns"
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        k = short(r["Kernel Name"])
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# per-kernel device time from {path} (ncu --metrics gpu__time_duration.sum, cold-cache, serialised)\n")
        f.write(f"# total {total:.0f} {unit} over {sum(cnt.values())} launches\n")
        f.write(f"{'share':>7} {'count':>6} {'total':>14} {'avg':>12}  kernel\n")
        for k in sorted(tot, key=tot.get, reverse=True):
            f.write(f"{100 * tot[k] / total:6.2f}% {cnt[k]:6d} {tot[k]:14.0f} {tot[k] / cnt[k]:12.1f}  {k}\n")
    print(open(out).read())


def report(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# key metrics from {path} (ncu --set full --clock-control none)\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n## {short(d['Kernel Name'])}  grid={d['Grid Size']} block={d['Block Size']}\n")
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write(f"{h:75s} {r[i]:>18s} {units[i]}\n")
    print(open(out).read())


def traffic(path, out, algorithmic=None, note=""):
    import json
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def val(d, key):
        v = float(d[key].replace(",", ""))
        u = units[hdr.index(key)]
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)

    recs = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rd, wr = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
        alg = None
        if algorithmic:
            if str(algorithmic).startswith("xCTA:"):            # algorithmic bytes = CTAs of the launch x bytes per CTA (one tile each)
                ctas = 1
                for v in re.findall(r"\d+", d["Grid Size"]):
                    ctas *= int(v)
                alg = ctas * float(str(algorithmic)[5:])
            else:
                alg = float(algorithmic)
        recs.append({"kernel": short(d["Kernel Name"]), "grid": d["Grid Size"], "block": d["Block Size"],
                     "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": rd + wr,
                     "duration_us_under_ncu": val(d, "gpu__time_duration.sum"),
                     "algorithmic_bytes": alg,
                     "launch": note, "source": path})
    with open(out, "w") as f:
        json.dump(recs, f, indent=1)
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
