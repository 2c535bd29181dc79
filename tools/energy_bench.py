"""Energy per kernel (run on the GPU box): every stage of the encoder step is looped ALONE for ~2 s while NVML's total-energy
counter runs, next to cuBLAS on the same GEMM shapes and on 8192^3 (the shape MEASURED_PEAKS.json quotes).  Output: one JSON
with W, J per launch, TFLOP/s or GB/s, pJ per FLOP / per byte and the SM clock under load of each loop, plus the sum over one
c2 step's launch counts against the step's own measured joules (bench.py `roofline.energy`).  VERDICT r1 next-2: the step
runs under the 1 kW power cap -- is it made of kernels at their own energy floor?

    python tools/energy_bench.py [--seconds 2.0] [--out gpurun_out/energy.json]
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
import torch

from sasvqa_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=2.0)
ap.add_argument("--chunk-frames", type=int, default=2048)
ap.add_argument("--out", default="gpurun_out/energy.json")
args = ap.parse_args()

torch.cuda.set_device(0)
pynvml.nvmlInit()
vis = os.environ.get("CUDA_VISIBLE_DEVICES")
H = pynvml.nvmlDeviceGetHandleByIndex(int(vis.split(",")[0]) if vis else 0)
M = args.chunk_frames * 197


class Clock(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.s, self.p, self.stop_evt = [], [], threading.Event()

    def run(self):
        while not self.stop_evt.is_set():
            self.s.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
            self.p.append(pynvml.nvmlDeviceGetPowerUsage(H) / 1e3)
            self.stop_evt.wait(0.05)


def measure(name, fn, work, unit, seconds=None):
    """loops fn for ~seconds; work = FLOP (unit 'flop') or bytes (unit 'byte') per launch"""
    seconds = seconds or args.seconds
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 10 / 1e3
    n = max(20, int(seconds / per))
    t_end = time.time() + 0.5                                 # settle clocks under this kernel's own load first
    while time.time() < t_end:
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
    clk = Clock()
    clk.start()
    j0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(H)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    j1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(H)
    clk.stop_evt.set()
    clk.join()
    ms = e0.elapsed_time(e1)
    joules = (j1 - j0) / 1e3
    rec = {"launches": n, "ms_per_launch": ms / n, "watts": joules / (ms / 1e3), "joules_per_launch": joules / n,
           "sm_mhz": statistics.median(clk.s) if clk.s else None, "power_w_sampled_max": max(clk.p) if clk.p else None}
    if unit == "flop":
        rec["tflops"] = work / (ms / n / 1e3) / 1e12
        rec["pj_per_flop"] = joules / n / work * 1e12
    else:
        rec["gbs"] = work / (ms / n / 1e3) / 1e9
        rec["pj_per_byte"] = joules / n / work * 1e12
    print(name, json.dumps(rec), flush=True)
    return rec


out = {"chunk_frames": args.chunk_frames, "rows": M, "seconds_per_loop": args.seconds, "kernels": {}, "cublas": {}}
# idle power (context up, nothing running)
time.sleep(1.0)
j0, t0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(H), time.time()
time.sleep(1.5)
out["idle_watts"] = (pynvml.nvmlDeviceGetTotalEnergyConsumption(H) - j0) / 1e3 / (time.time() - t0)
print("idle W", out["idle_watts"], flush=True)

shapes = [("gemm_qkv", 2304, 768, 0), ("gemm_out_proj", 768, 768, 2), ("gemm_fc1", 3072, 768, 1), ("gemm_fc2", 768, 3072, 2)]
for name, N, K, mode in shapes:
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    vec = torch.randn(N, device="cuda")
    x = torch.zeros(M, N, device="cuda") if mode == 2 else None
    out["kernels"][name] = measure(name, lambda: ops.test_gemm(a, b, mode, vec, out_f32=x), 2.0 * M * N * K, "flop")
    c = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    out["cublas"][name] = measure("cublas_" + name, lambda: torch.matmul(a, b.t(), out=c), 2.0 * M * N * K, "flop")
    del a, b, x, c
a = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
c = torch.empty(8192, 8192, dtype=torch.bfloat16, device="cuda")
out["cublas"]["8192^3"] = measure("cublas_8192^3", lambda: torch.matmul(a, b, out=c), 2.0 * 8192 ** 3, "flop")
del a, b, c

qkv = torch.randn(M, 2304, device="cuda").to(torch.bfloat16)
out["kernels"]["attention"] = measure("attention", lambda: ops.test_attention(qkv), 12 * 4.0 * 197 * 197 * 64 * args.chunk_frames,
                                      "flop")
out["kernels"]["attention"]["algorithmic_gbs"] = M * (2304 + 768) * 2 / (out["kernels"]["attention"]["ms_per_launch"] / 1e3) / 1e9
del qkv
x = torch.randn(M, 768, device="cuda")
g, bt = torch.ones(768, device="cuda"), torch.zeros(768, device="cuda")
out["kernels"]["layernorm"] = measure("layernorm", lambda: ops.test_layernorm(x, g, bt), M * (768 * 4 + 768 * 2), "byte")
del x

# ---- one c2 step (256 clips x 128 frames = 16 chunks of 2048 frames x 12 layers) out of these kernels
per_step = {"gemm_qkv": 192, "gemm_out_proj": 192, "gemm_fc1": 192, "gemm_fc2": 192, "attention": 192, "layernorm": 384}
out["c2_step_from_isolated_kernels"] = {
    "launches_per_step": per_step,
    "joules": {k: out["kernels"][k]["joules_per_launch"] * n for k, n in per_step.items()},
    "ms": {k: out["kernels"][k]["ms_per_launch"] * n for k, n in per_step.items()},
}
out["c2_step_from_isolated_kernels"]["joules_total"] = sum(out["c2_step_from_isolated_kernels"]["joules"].values())
out["c2_step_from_isolated_kernels"]["ms_total"] = sum(out["c2_step_from_isolated_kernels"]["ms"].values())
gemm_flop = sum(2.0 * M * N * K * per_step[n] for n, N, K, _ in shapes)
out["ours_gemm_set_pj_per_flop"] = sum(out["c2_step_from_isolated_kernels"]["joules"][n] for n, *_ in shapes) / gemm_flop * 1e12
out["cublas_gemm_set_pj_per_flop"] = sum(out["cublas"][n]["joules_per_launch"] * per_step[n] for n, *_ in shapes) / gemm_flop * 1e12
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump(out, open(args.out, "w"), indent=1)
print("wrote", args.out)
