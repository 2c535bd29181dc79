"""Summarises ncu exports under gpurun_out/ into small text files for profiles/ (run here, no GPU)."""
import collections, csv, json, subprocess, sys

def launch_list(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("sasvqa::<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v * 1e3 if unit == "s" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'share':>7s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k[-58:]:58s} {v[0]:8d} {v[1]:10.3f} {100 * v[1] / tot:6.1f}%")
    out.append(f"{'TOTAL':58s} {sum(v[0] for v in agg.values()):8d} {tot:10.3f}")
    return "\n".join(out)

def raw_metrics(rep, wanted):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
        out.append({w: d[w] for w in wanted if w in d})
    return out

if __name__ == "__main__":
    what = sys.argv[1]
    if what == "list":
        print(launch_list(sys.argv[2]))
    else:
        wanted = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                  "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                  "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
                  "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]
        for m in raw_metrics(sys.argv[2], wanted):
            for k, (v, u) in m.items():
                print(f"{k} [{u}] = {v[-110:]}")
            print("----")
