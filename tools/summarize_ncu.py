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

def gemm_json(rep, out_path, chunk_frames, capture):
    """profiles/rNN/ncu_gemm.json: DRAM bytes and tensor-pipe % per per-layer GEMM mode + the hash of the GEMM sources the
    capture was taken on (bench.py refuses the numbers when the sources have changed since)."""
    import hashlib, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()
    for name in ("gemm_tcgen05.cu",):
        h.update(open(os.path.join(root, "sas-vqa_b200", "csrc", name), "rb").read())
    wanted = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    def num(m, key, scale_units=True):
        v, u = m[key]
        x = float(v.replace(",", ""))
        if scale_units:
            x *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "%": 1.0}.get(u, 1.0)
        return x
    rows = []
    for m in raw_metrics(rep, wanted):
        name = m["Kernel Name"][0]
        mode = int(name.split("gemm_tcgen05_kernel<")[1].split(">")[0])
        rows.append(dict(mode=mode, ms=num(m, "gpu__time_duration.sum"),
                         dram_bytes=num(m, "dram__bytes_read.sum") + num(m, "dram__bytes_write.sum"),
                         dram_read=num(m, "dram__bytes_read.sum"), dram_write=num(m, "dram__bytes_write.sum"),
                         tensor_pipe_pct=num(m, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                         dram_pct=num(m, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
    modes = {}
    resid = sorted([r for r in rows if r["mode"] == 2], key=lambda r: r["ms"])
    for r in rows:
        if r["mode"] == 0: modes["qkv"] = r
        if r["mode"] == 1: modes["fc1"] = r
    if resid:
        modes["out_proj"], modes["fc2"] = resid[0], resid[-1]
    assert set(modes) == {"qkv", "out_proj", "fc1", "fc2"}, modes.keys()
    out = {"source_sha256_16": h.hexdigest()[:16], "sources": ["gemm_tcgen05.cu"], "chunk_frames": chunk_frames,
           "capture": capture, "modes": modes,
           "mean_dram_bytes_per_launch": sum(m["dram_bytes"] for m in modes.values()) / 4,
           "note": "one launch per mode from `ncu --set full --clock-control none` (cold-cache, serialised: shares agree with the live "
                   "CUDA-event timing, absolutes do not)"}
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "list":
        print(launch_list(sys.argv[2]))
    elif what == "gemm-json":
        gemm_json(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5])
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
