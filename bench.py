#!/usr/bin/env python
"""bench.py — image-pairs/sec of the MicKey inference hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE configs[1] — one synthetic 720x540 pair per step and per GPU,
DINOv2 ViT-S/14 backbone, 512 hypotheses (IT_MATCHES 8 x IT_RANSAC 64), 2048 sampled matches, seeded
random-init weights (mickey_b200.weights.synthetic_state_dict).  A step = one model(data) call =
extraction of both images + dual-softmax matching + RANSAC pose.  Pairs are independent, so N GPUs
run N pairs per step (weak scaling) and exchange ONE all-gather of the packed [B,13] poses per step.

One JSON line on stdout (rank 0).  `value` is measured with inputs resident in HBM; `e2e` goes through
the same public API with pinned-host inputs (H2D of both images and D2H of the pose inside the timed
region).  `roofline*` objects come from CUDA-event timings of each kernel class taken live in extra
profiled steps of this run (mk_profile_*), divided into the algorithmic work stated in DESIGN.md.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mickey_b200.config import mickey_cfg  # noqa: E402
from mickey_b200.weights import synthetic_checkpoint, synthetic_state_dict  # noqa: E402

H_IMG, W_IMG = 720, 540
VARIANT, IT_MATCHES, IT_RANSAC = "vits", 8, 64
PAIRS_PER_STEP = 1
WORKLOADS = {
    # name: (variant, it_matches, it_ransac, pairs per GPU per step, description)
    "c2": ("vits", 8, 64, 1, "BASELINE configs[1]: single 720x540 synthetic pair, ViT-S/14, 512 hypotheses (8x64), 2048 sampled matches"),
    "c3": ("vitb", 16, 64, 32, "BASELINE configs[2]: batch of 32 synthetic 720x540 pairs, ViT-B/14, 1024 hypotheses (16x64), 2048 sampled matches"),
}
VIT_DIMS = {"vits": (384, 12), "vitb": (768, 12), "vitl": (1024, 24)}
K_TOY = [[549.7, 0.0, 268.7], [0.0, 549.7, 351.8], [0.0, 0.0, 1.0]]
WORKLOAD = "BASELINE configs[1]: single 720x540 synthetic pair, ViT-S/14, 512 hypotheses (8x64), 2048 sampled matches"


def synthetic_pair(batch, seed):
    g = torch.Generator().manual_seed(seed)
    im0 = torch.rand(batch, 3, H_IMG, W_IMG, generator=g)
    im1 = torch.rand(batch, 3, H_IMG, W_IMG, generator=g)
    K = torch.tensor(K_TOY)[None].repeat(batch, 1, 1)
    return im0, im1, K


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "tflops_burst": p["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json; sustained bf16 GEMM, copy bandwidth)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------------------------------------------
# algorithmic work per pair
# ---------------------------------------------------------------------------------------------------------------
def work_model(D=384, depth=12, n_pairs=1):
    gh, gw = H_IMG // 14, W_IMG // 14
    N, T = gh * gw, gh * gw + 1
    imgs = 2 * n_pairs
    M = imgs * T
    flops = {
        "vit.patch_embed": 2 * imgs * N * 588 * D,
        "vit.qkv": depth * 2 * M * D * 3 * D, "vit.proj": depth * 2 * M * D * D,
        "vit.fc1": depth * 2 * M * D * 4 * D, "vit.fc2": depth * 2 * M * D * 4 * D,
        "vit.attention": depth * imgs * 4 * T * T * D,
    }
    px = imgs * N
    dims = [D, 512, 256, 128]
    c3 = c1 = 0
    for r in range(3):
        cin, cout = dims[r], dims[r + 1]
        c3 += 4 * 2 * px * 9 * (cin * cout + cout * cout)
        c1 += 4 * 2 * px * cin * cout
    c3 += 3 * 2 * px * 9 * (128 * 64 + 64 * 64) + 2 * px * 9 * (128 * 128 + 128 * 128)
    c1 += 3 * 2 * px * 128 * 64
    flops["head.conv3x3"], flops["head.conv1x1"] = c3, c1
    flops["head.att.qkv"] = 3 * 4 * 2 * px * 128 * 384
    flops["head.att.merge_ln"] = 3 * 4 * 2 * px * 128 * 128
    flops["head.att.mlp0"] = 3 * 4 * 2 * px * 256 * 256
    flops["head.att.mlp2_ln"] = 3 * 4 * 2 * px * 256 * 128
    nbytes = {
        # dual-softmax pass 2: read both descriptor sets + scores, write scores, kp_scores, final_scores
        "match.dual_softmax": n_pairs * (2 * 128 * N * 4 + 2 * N * 4 + 3 * N * N * 4),
        # solver: final_scores read once + keypoints/depth, 52 B out  (SURVEY.md §8d "RANSAC bytes")
        "solve.sample_outer": n_pairs * (N * N * 4),
    }
    return flops, nbytes


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled through NVML DURING the timed region (the same counters
    `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.power, self.reasons, self.stop_flag, self.max_sm, self.err = index, [], [], set(), False, None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            while not self.stop_flag:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:          # noqa: BLE001
            self.err = repr(e)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        sm = sorted(self.sm)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
               "samples": len(sm), "power_w_max": max(self.power) if self.power else None}
        if self.err:
            out["error"] = self.err
        return out


def usable_cpus():
    """Host threads this process may really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:      # noqa: BLE001
        pass
    return n


# ---------------------------------------------------------------------------------------------------------------
def cpu_threads():
    # torch CPU kernels at these sizes stop scaling (and then regress badly) past a few dozen threads
    return min(usable_cpus(), 32)


def cpu_reference_run(n_timed, n_warm, threads, budget_s=25.0):
    """The reference's CPU path restated by the oracle (kind 'port': the Python reference cannot travel to
    the GPU box; oracle/mickey_oracle.py is pinned to it by tests/golden), fp32.  Returns per-pair seconds."""
    from oracle import mickey_oracle as mo
    torch.set_num_threads(threads)
    cfg = mickey_cfg(VARIANT, IT_MATCHES, IT_RANSAC, float16=False)
    sd = synthetic_state_dict(cfg, seed=0)
    im0, im1, K = synthetic_pair(1, seed=0)
    durs, spent = [], 0.0
    with torch.no_grad():
        for i in range(n_warm + n_timed):
            data = {"image0": im0, "image1": im1, "K_color0": K, "K_color1": K}
            torch.manual_seed(i)
            t0 = time.perf_counter()
            mo.model_forward(sd, data, cfg)
            dt = time.perf_counter() - t0
            spent += dt
            if i >= n_warm:
                durs.append(dt)
            if spent > budget_s and durs:
                break
    return durs


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads()
    # a "step" is one pair through the CPU path; warm-up steps are run but not timed; the run is bounded
    steps, warm = max(1, min(args.steps, 8)), max(1, min(args.warmup, 1))
    durs = cpu_reference_run(steps, warm, threads, budget_s=120.0)
    steps = len(durs)
    val = len(durs) / sum(durs)
    line = {"impl": "reference", "metric": "image-pairs/sec @720x540", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * sum(durs) / len(durs), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference CPU path (oracle port of the PyTorch reference), rank 0 only"},
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port",
                             "sample": f"{steps} pairs of the bench workload (steps clamped to 8 / 120 s, 1 warm-up), "
                                       f"torch threads = {threads} of {usable_cpus()} usable"},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS),
                    help="c2 (default, the configuration the metric is quoted on) or c3 (B=32, ViT-B; extra data point)")
    ap.add_argument("--depth", type=int, default=int(os.environ.get("MICKEY_PIPELINE_DEPTH", "3")),
                    help="steps kept in flight on alternating engines/streams (1 = strictly one step at a time)")
    args = ap.parse_args()
    global VARIANT, IT_MATCHES, IT_RANSAC, WORKLOAD, PAIRS_PER_STEP
    VARIANT, IT_MATCHES, IT_RANSAC, PAIRS_PER_STEP, WORKLOAD = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from mickey_b200.model import build_model
    from mickey_b200 import dist as mkdist

    cfg = mickey_cfg(VARIANT, IT_MATCHES, IT_RANSAC)
    model = build_model(cfg, synthetic_checkpoint(cfg, seed=0, with_backbone=True))
    model.static_outputs = True     # hand out the engine's static output buffers (no per-call clones)
    B = PAIRS_PER_STEP
    im0, im1, K = synthetic_pair(B, seed=rank)
    dev_data = {"image0": im0.to(dev), "image1": im1.to(dev), "K_color0": K.to(dev), "K_color1": K.to(dev)}
    pin = {"image0": im0.pin_memory(), "image1": im1.pin_memory()}
    pose_host = torch.empty(B, 13).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step_device():
        data = dict(dev_data)
        R, t = model(data)
        packed = torch.cat([R.reshape(B, 9), t.reshape(B, 3), data["inliers"].reshape(B, 1)], dim=1)
        return mkdist.gather_poses(packed)

    def step_e2e():
        # pinned HOST images go straight into model(): the engine's H2D copies land in its static input buffer
        data = {"image0": pin["image0"], "image1": pin["image1"], "K_color0": dev_data["K_color0"], "K_color1": dev_data["K_color1"]}
        R, t = model(data)
        packed = torch.cat([R.reshape(B, 9), t.reshape(B, 3), data["inliers"].reshape(B, 1)], dim=1)
        allp = mkdist.gather_poses(packed)
        pose_host.copy_(allp[rank * B:(rank + 1) * B], non_blocking=True)
        return allp

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, flush_l2):
        """K steps between two synchronisation points.  flush_l2=True: one event pair per step with a 256 MiB write in
        between (steps strictly one after the other).  flush_l2=False (pipelined mode: two steps in flight, so there is
        no gap to flush in): one event pair around all K steps; the per-step working set exceeds the L2 (see config)."""
        barrier()
        if flush_l2:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for e0, e1 in evs:
                flush.fill_(1)
                e0.record()
                fn()
                e1.record()
            barrier()
            ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)       # max over ranks
        return float(t.item())

    torch.manual_seed(1234 + rank)
    # (1) latency: one step at a time (pipeline depth 1), L2 flushed between steps
    model.pipeline_depth = 1
    for _ in range(args.warmup):
        step_device()
    ncu_range = os.environ.get("MICKEY_NCU_RANGE") == "1"     # `ncu --profile-from-start off`: capture the latency steps only
    if ncu_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    latency_ms = timed(step_device, args.steps, True) / args.steps
    if ncu_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    eng = model._engine()
    # (2) throughput: two steps in flight on alternating engines/streams (inputs are constant device tensors)
    model.pipeline_depth = max(1, args.depth)
    model.assume_inputs_ready = True
    for _ in range(max(args.warmup, 4 * max(1, args.depth))):        # each of the 2 engines x 2 buffer sets: one eager call + one capture
        step_device()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    engines = model._engine_pool()
    l0 = sum(e.total_kernel_launches for e in engines)
    total_ms = timed(step_device, args.steps, False)
    launches = sum(e.total_kernel_launches for e in engines) - l0
    clock_info = clocks.finish() if rank == 0 else None
    ws_bytes = sum(int(e.ws.numel()) for e in engines)

    for _ in range(4 * max(1, args.depth)):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps, False)
    model.pipeline_depth = 1
    model.assume_inputs_ready = False

    # ---- per-kernel-class device times (CUDA events on the launch stream), extra profiled steps
    prof = {}
    if rank == 0:
        n_prof = 3
        eng.profile(True)
        model.use_graph = False                 # per-kernel events need eager launches
        for _ in range(n_prof):
            flush.fill_(1)
            model(dict(dev_data))               # no collective here: the other ranks are already done
        model.use_graph = True
        raw = eng.profile_read()
        eng.profile(False)
        prof = {k: {"scopes_per_step": v[0] / n_prof, "ms_per_step": v[1] / n_prof} for k, v in raw.items()}

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    value = world * B * args.steps / (total_ms / 1e3)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    flops, nbytes = work_model(D=VIT_DIMS[VARIANT][0], depth=VIT_DIMS[VARIANT][1], n_pairs=B)
    nbytes["solve.sample_outer"] = B * (H_IMG // 14 * (W_IMG // 14)) ** 2 * 4

    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:      # noqa: BLE001
        traffic = {}

    def roof(cls):
        r = roof_inner(cls)
        if r is not None:
            r["traffic"] = traffic.get(cls)       # DRAM bytes per launch from the committed ncu --set full capture
        return r

    def roof_inner(cls):
        if cls not in prof or prof[cls]["ms_per_step"] <= 0:
            return None
        ms, n = prof[cls]["ms_per_step"], max(prof[cls]["scopes_per_step"], 1)
        if cls in flops:
            ach = flops[cls] / (ms / 1e3) / 1e12
            return {"kernel": cls, "bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                    "frac": ach / peaks["tflops"], "traffic": None, "launches_per_step": n, "avg_launch_ms": ms / n,
                    "peak_source": peaks["source"]}
        if cls in nbytes:
            ach = nbytes[cls] / (ms / 1e3) / 1e9
            return {"kernel": cls, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": None, "launches_per_step": n, "avg_launch_ms": ms / n,
                    "peak_source": peaks["source"]}
        return None

    dominant = max(prof, key=lambda k: prof[k]["ms_per_step"]) if prof else None

    def attention_roof():
        """The tensor roofline object plus the bound that head_dim 64 really imposes: one MUFU.EX2 per logit at 16 per
        clock per SM (measured, tools/ubench/pipe_rates.cu) costs twice the tile's MMA time."""
        r = roof("vit.attention")
        if r is None:
            return None
        D_, depth_ = VIT_DIMS[VARIANT]
        T_ = (H_IMG // 14) * (W_IMG // 14) + 1
        exps = depth_ * 2 * B * (D_ // 64) * T_ * (-(-T_ // 128) * 128)           # padded key tiles are exponentiated too
        sm_mhz = (clock_info or {}).get("sm_mhz") or 1965.0
        n_sm = torch.cuda.get_device_properties(0).multi_processor_count
        rate = exps / (prof["vit.attention"]["ms_per_step"] / 1e3) / (sm_mhz * 1e6) / n_sm
        r["mufu"] = {"achieved": rate, "peak": 16.0, "unit": "exp2/clk/SM", "frac": rate / 16.0,
                     "note": "head_dim 64: 1024 MUFU clk vs 512 tensor clk per 128x128 tile, so 16 exp2/clk/SM caps a kernel that sends "
                             "every exp2 to the MUFU at half of the tensor peak; this one evaluates a quarter of them on the FMA pipe"}
        return r
    vit_gemm_ms = sum(prof[k]["ms_per_step"] for k in ("vit.qkv", "vit.proj", "vit.fc1", "vit.fc2") if k in prof)
    vit_gemm_fl = sum(flops[k] for k in ("vit.qkv", "vit.proj", "vit.fc1", "vit.fc2"))
    line = {
        "metric": "image-pairs/sec @720x540", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "latency_ms_single_step": latency_ms,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (tensor core), f32 matcher+solver", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": B, "parallelism": f"dp{world} (pairs sharded, one all-gather of [B,13] poses)",
                   "pipelining": f"{max(1, args.depth)} steps in flight (engines on alternating CUDA streams); latency_ms_single_step is the "
                                 "un-pipelined time of one step with the L2 flushed in between",
                   "l2": f"no flush in pipelined mode: each step streams its {ws_bytes // max(1, args.depth) / 1e6:.0f} MB workspace + 46 MB of "
                         "N x N outputs + 125 MB of weights, larger than the 126 MB L2; the latency run flushes a 256 MiB buffer",
                   "weights": "seeded random init", "launch": "one mk_forward C call per step, replayed from a CUDA graph"},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(2 * B * 3 * H_IMG * W_IMG * 4), "d2h_bytes_per_step": int(B * 13 * 4)},
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": (attention_roof() if dominant == "vit.attention" else roof(dominant)) if dominant else None,
        "roofline_vit_gemm": ({"kernel": "vit.qkv+proj+fc1+fc2", "bound": "tensor", "achieved": vit_gemm_fl / (vit_gemm_ms / 1e3) / 1e12,
                               "peak": peaks["tflops"], "unit": "TFLOP/s",
                               "frac": vit_gemm_fl / (vit_gemm_ms / 1e3) / 1e12 / peaks["tflops"], "traffic": None}
                              if vit_gemm_ms > 0 else None),
        "roofline_attention": attention_roof(), "roofline_head_conv": roof("head.conv3x3"),
        "roofline_matcher": roof("match.dual_softmax"), "roofline_sampler": roof("solve.sample_outer"),
        "stage_ms": {k: round(v["ms_per_step"], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms_per_step"])},
    }
    if world == 1 and not args.no_cpu_baseline and args.workload == "c2":
        cores = cpu_threads()
        durs = cpu_reference_run(n_timed=3, n_warm=1, threads=cores, budget_s=25.0)
        line["cpu_baseline"] = {"value": len(durs) / sum(durs), "unit": "pairs/s", "cores": cores, "kind": "port",
                                "sample": f"{len(durs)} pair(s) of the same workload after 1 warm-up ({sum(durs) / len(durs):.2f} s/pair), "
                                          f"fp32 oracle, torch threads = {cores} of {usable_cpus()} usable"}
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
