#!/usr/bin/env python
"""bench.py -- graph-NCA cell-updates/s on B200 (BASELINE.json metric), one JSON line on stdout.

Workload (default, BASELINE.json configs[1]): graph-augmented NCA forward rollout, 16 channels, 40x40, batch 8,
96 steps, fire_rate 0.5, torus graph shift, 8 of 72 offsets per step, trained weights (tests/golden fixture),
growth from the single-cell seed.  A "step" of this bench = ONE such rollout (B*T*H*W = 1,228,800 cell-updates).

  value : rollouts timed with CUDA events, x0 / weights / schedule resident in HBM, L2 flushed between iterations.
  e2e   : the same through the public API with HOST buffers: pinned x0 -> H2D, schedule draws (random.sample per
          step, as T forward calls would), rollout, x_T -> D2H.  Wall clock around each call (sync both sides).
  roofline      : dominant kernel, CUDA events around its launches (library hook), fp32-FMA bound.
  cpu_baseline  : oracle/ port of the reference's PyTorch path on the host cores (N=1, rank 0).
  --impl reference : the oracle port alone, same JSON shape.

Other workloads: --workload c1 (classic NCA rollout), c3 (training step fwd+bwd, B=32, T=64: also part of the default
line as `fwd_bwd`), c4 (c3 with an in-kernel damage mask),
c3l (the same at T=300, long regime), c5s (256x256x32 scale-up slice, streaming kernels), c5 (the full per-GPU share of
BASELINE configs[4]: B=128, T=1000, damage at t=500 -- seconds per rollout: run it with --steps 2 --warmup 1).
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# dense algorithmic work per cell-update (SURVEY 8d / BASELINE.md 4), graph model C=16 hidden=128 k=8
FLOP_FWD_GRAPH_S = 17.9e3
FLOP_FWDBWD_GRAPH_S = 53.6e3
FLOP_FWD_GRAPH_L = 36.7e3
FP32_LANES = 148 * 128 * 2           # FMA lanes x 2 flop
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/)
NCU_DRAM_BYTES = {("c2", "k_rep_fwd"): 1090816,     # profiles/r01_rep_fwd_summary.md: 1.090304 MB read, 0 written (x_T still in L2)
                  ("c5", "k_update"): 112492544}    # profiles/r01_k_update_summary.md (B=16 slice): 81.27 MB read + 31.22 MB written
# sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active of the same captures: the REAL pipe utilisation next to the
# dense-equivalent fraction (which counts the flops of the cells the kernels skip)
NCU_FMA_PIPE_PCT = {("c2", "k_rep_fwd"): 18.5, ("c3", "k_rep_bwd"): 30.1, ("c5", "k_update"): 31.8}


def load_weights(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name)).items()}


def workload_cfg(name):
    if name == "c2":
        return dict(name="c2: graph NCA fwd rollout, seed growth", C=16, H=40, W=40, B=8, T=96, hidden=128,
                    fire_rate=0.5, message_every=1, train=False, flop=FLOP_FWD_GRAPH_S)
    if name == "c3":
        return dict(name="c3: graph NCA training step (fwd+bwd, short regime)", C=16, H=40, W=40, B=32, T=64,
                    hidden=128, fire_rate=0.7, message_every=3, train=True, flop=FLOP_FWDBWD_GRAPH_S)
    if name == "c1":
        return dict(name="c1: classic NCA fwd rollout, seed growth (the reference's CPU-runnable case, here on the GPU)",
                    C=16, H=40, W=40, B=8, T=96, hidden=128, fire_rate=0.5, message_every=1, train=False,
                    flop=17.0e3, classic=True)
    if name == "c4":
        return dict(name="c4: graph NCA training step with damage (fwd+bwd, short regime, damaged batch)", C=16, H=40, W=40,
                    B=32, T=64, hidden=128, fire_rate=0.7, message_every=3, train=True, flop=FLOP_FWDBWD_GRAPH_S,
                    damage=True)
    if name == "c3l":
        return dict(name="c3l: graph NCA training step (fwd+bwd, long regime)", C=16, H=40, W=40, B=32, T=300,
                    hidden=128, fire_rate=0.7, message_every=3, train=True, flop=FLOP_FWDBWD_GRAPH_S)
    if name == "c5s":
        return dict(name="c5 slice: 256x256x32 fwd rollout", C=32, H=256, W=256, B=16, T=20, hidden=128,
                    fire_rate=0.5, message_every=1, train=False, flop=FLOP_FWD_GRAPH_L)
    if name == "c5":      # BASELINE configs[4] per GPU: batch 1024 over 8 GPUs = 128 per GPU, 1000 steps, damage at t = 500
        return dict(name="c5: 256x256x32 regeneration rollout, B=128 per GPU, T=1000, circle damage at t=500", C=32, H=256,
                    W=256, B=128, T=1000, hidden=128, fire_rate=0.5, message_every=1, train=False, flop=FLOP_FWD_GRAPH_L,
                    damage=True, damage_step=500, damage_size=64)
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons while the timed region runs (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons = index, False, [], set()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)       # the default timed region is ~20 ms: sample densely enough to see it

    def result(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_oracle_inputs(cfg, seed=42):
    """Deterministic draws for the CPU port: offsets per step + fire uniforms."""
    from oracle import nca_oracle as O
    random.seed(seed)
    torch.manual_seed(seed)
    offs = O.build_offsets(4)
    chosens = [random.sample(offs, 8) for _ in range(cfg["T"])]
    return chosens


def cpu_port_rollout(cfg, params, T, chosens):
    """The reference's PyTorch path restated (oracle/), CPU, no grad: T forward calls from the seed."""
    from oracle import nca_oracle as O
    oc = O.StepConfig(update_gain=0.05, alpha_thr=0.12, graph=True, message_gain=0.25, hidden_only=True,
                      zero_padded_shift=False)
    x = O.make_seed(cfg["C"], cfg["H"], cfg["B"])
    with torch.no_grad():
        for t in range(T):
            fu = torch.rand(cfg["B"], 1, cfg["H"], cfg["W"])
            c = oc
            if cfg["message_every"] > 1 and t % cfg["message_every"] != 0:
                c = O.StepConfig(**{**oc.__dict__, "message_gain": 0.0})
            x = O.nca_step(x, params, c, cfg["fire_rate"], fu, chosens[t])
    return x


def time_cpu_port(cfg, reps=2):
    params = load_weights("weights_graph_ep960.npz")
    if cfg["C"] != 16:
        return None
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    chosens = build_oracle_inputs(cfg)
    T = cfg["T"]
    cpu_port_rollout(cfg, params, min(T, 8), chosens)           # warm-up
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_port_rollout(cfg, params, T, chosens)
        best = min(best, time.perf_counter() - t0)
    updates = cfg["B"] * T * cfg["H"] * cfg["W"]
    return {"value": updates / best, "unit": "cell-updates/s", "cores": threads, "kind": "port",
            "sample": f"full workload: B={cfg['B']} T={T} {cfg['H']}x{cfg['W']}x{cfg['C']} forward rollout, best of {reps}, "
                      f"{best:.2f} s per rollout, torch CPU {torch.__version__}"}


def run_reference_arm(args, cfg, rank):
    if rank != 0:
        return
    params = load_weights("weights_graph_ep960.npz")
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    chosens = build_oracle_inputs(cfg)
    T = cfg["T"]
    for _ in range(args.warmup):
        cpu_port_rollout(cfg, params, T, chosens)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_rollout(cfg, params, T, chosens)
    dt = time.perf_counter() - t0
    updates = cfg["B"] * T * cfg["H"] * cfg["W"]
    val = updates * args.steps / dt
    line = {"impl": "reference", "metric": "graph-NCA cell-updates/s (fwd)", "value": val, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "B": cfg["B"], "T": T, "grid": [cfg["H"], cfg["W"]], "channels": cfg["C"],
                       "note": "reference's PyTorch CPU path restated in oracle/ (the reference checkout does not travel)"},
            "cpu_baseline": {"value": val, "unit": "cell-updates/s", "cores": threads, "kind": "port",
                             "sample": "full workload per step"},
            "e2e": {"value": val, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_workload(cfg, args, world, rank, local_rank, dev, steps, warmup, with_roofline=True):
    """Times one workload: returns dict(value, ms_per_step, e2e, launches, roofline, clocks, data)."""
    import torch.distributed as dist
    import graph_neural_cellular_automata_b200 as G
    from graph_neural_cellular_automata_b200 import _lib
    from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
    from graph_neural_cellular_automata_b200.utils.nca_init import make_seed
    import ctypes

    lib = _lib.load()
    C_, H, W, B, T = cfg["C"], cfg["H"], cfg["W"], cfg["B"], cfg["T"]
    torch.manual_seed(42 + rank); random.seed(42 + rank)
    if cfg.get("classic"):
        model = G.NeuralCA(C_, update_hidden=cfg["hidden"], img_size=H, update_gain=0.05, alpha_thr=0.12)
        model.load_state_dict(load_weights("weights_classic_ep990.npz"), strict=False)
        data = "synthetic (seed growth; trained classic 40x40 gecko weights shipped as a fixture)"
    else:
        model = G.NeuralCAGraph(C_, update_hidden=cfg["hidden"], img_size=H, update_gain=0.05, alpha_thr=0.12,
                                message_gain=0.25, hidden_only=True, graph_zero_padded_shift=False)
    if cfg.get("classic"):
        pass
    elif C_ == 16:
        model.load_state_dict(load_weights("weights_graph_ep960.npz"), strict=False)
        data = "synthetic (seed growth; trained 40x40 gecko weights shipped as a fixture)"
    else:
        with torch.no_grad():
            model.update_net[2].weight.normal_(0, 0.05)
        data = "synthetic (seeded random weights, zero-mean W2 sigma 0.05)"
    model = model.to(dev)
    if cfg["train"]:
        target = torch.from_numpy(np.load(os.path.join(GOLDEN, "target_gecko_surrogate.npy"))).to(dev)

    def make_x0():
        if C_ == 16 and cfg["train"]:       # pool-like states: seeds grown for a random number of steps
            x = make_seed(C_, H, B, device=dev)
            with torch.no_grad():
                for t in range(48):
                    x = model(x, fire_rate=0.6)
            return x.cpu()
        if C_ == 16:
            return make_seed(C_, H, B, device="cpu")
        x = torch.rand(B, C_, H, W)
        yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        disk = (((yy - H / 2) ** 2 + (xx - W / 2) ** 2) < (0.3 * H) ** 2).float()
        return x * disk

    x0_host = make_x0().pin_memory()
    x0_dev = x0_host.to(dev)
    xT_host = torch.empty_like(x0_host).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    updates = B * T * H * W

    if cfg["train"]:
        from graph_neural_cellular_automata_b200.training.trainer import premult_loss
        from graph_neural_cellular_automata_b200.training.optim import FusedNormalizedAdam
        from graph_neural_cellular_automata_b200.rollout import rollout_fwd_raw, rollout_bwd_raw
        opt = FusedNormalizedAdam(model, lr=2e-4, weight_decay=1e-5, normalize=True)
        IMPLN = {"auto": 0, "streaming": 1, "resident": 2, "banded": 3}[args.rollout_impl]

    def one_rollout(x0, sched):
        if cfg["train"]:     # fwd (history) + premult loss + BPTT + grad normalise + Adam: one training iteration's hot path
            desc, packed = model.model_desc(), model.packed_weights()
            xT, hist = rollout_fwd_raw(desc, packed, x0, sched, history=True, impl=IMPLN, keep_x=False)
            per, gxT = premult_loss(xT, target, 1.0 / (B * world))
            _, gflat = rollout_bwd_raw(desc, packed, hist, sched, gxT, impl=IMPLN)
            if world > 1:
                dist.all_reduce(gflat)
            opt.step(gflat)
            return xT
        with torch.no_grad():
            return rollout(model, x0, sched, impl=args.rollout_impl)

    dmg = None
    if cfg.get("damage"):      # one damage kind / size for the batch, per-sample positions (utils/damage.py), applied in-kernel at step 0
        from graph_neural_cellular_automata_b200.utils.damage import circle_mask
        dmg = circle_mask(x0_dev, int(cfg.get("damage_size", 5))).expand_as(x0_dev).contiguous()

    def new_schedule(seed):
        return make_schedule(model, B, H, W, T, fire_rate=cfg["fire_rate"], message_every=cfg["message_every"],
                             fire="philox", seed=seed, damage=dmg, damage_step=int(cfg.get("damage_step", 0)))

    # ---------------- device-resident timing (value) ----------------
    scheds = [new_schedule(1000 + i) for i in range(warmup + steps)]
    torch.cuda.synchronize()
    for i in range(warmup):
        one_rollout(x0_dev, scheds[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.gnca_launch_count()
    evs = []
    torch.cuda.synchronize()
    for i in range(steps):
        flush.fill_(float(i))                        # L2 flush between timed iterations (outside the events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_rollout(x0_dev, scheds[warmup + i])
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    launches = lib.gnca_launch_count() - launches0
    sampler.stop_flag = True
    sampler.join()
    if world > 1:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = world * updates * steps / (dev_ms_max * 1e-3)

    # ---------------- end-to-end through the public API with host buffers ----------------
    def e2e_once(seed):
        x0 = x0_host.to(dev, non_blocking=True)
        sched = new_schedule(seed)
        xT = one_rollout(x0, sched)
        xT_host.copy_(xT, non_blocking=True)
        torch.cuda.synchronize()

    if os.environ.get("GNCA_E2E_DEBUG"):
        import time as _t
        for i in range(5):
            t_a = _t.perf_counter(); x0_ = x0_host.to(dev, non_blocking=True); t_b = _t.perf_counter()
            sc_ = new_schedule(7000 + i); t_c = _t.perf_counter(); y_ = one_rollout(x0_, sc_); t_d = _t.perf_counter()
            xT_host.copy_(y_, non_blocking=True); t_e = _t.perf_counter(); torch.cuda.synchronize(); t_f = _t.perf_counter()
            print("[e2e debug] h2d %.3f sched %.3f launch %.3f d2h-call %.3f sync %.3f total %.3f ms" % (
                (t_b - t_a) * 1e3, (t_c - t_b) * 1e3, (t_d - t_c) * 1e3, (t_e - t_d) * 1e3, (t_f - t_e) * 1e3, (t_f - t_a) * 1e3),
                file=sys.stderr)
    import gc
    for i in range(8 if updates < 1e8 else 1):          # second-long rollouts (c5) need no repeated host warm-up
        e2e_once(5000 + i)
    gc.collect()
    gc.freeze()          # the timed loop allocates a handful of small objects per call; do not rescan the rest of the heap
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        e2e_once(6000 + i)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * updates * steps / float(t.item())
    sched_bytes = T * 4 * 2 + T * 8 * 2
    e2e = {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": int(x0_host.numel() * 4 + sched_bytes),
           "d2h_bytes_per_step": int(xT_host.numel() * 4), "ms_per_step": float(t.item()) / steps * 1e3}

    # ---------------- roofline of the dominant kernel (library event hook), rank 0 ----------------
    roof = None
    if with_roofline:            # every rank runs these iterations (the training step holds a collective); rank 0 reports
        lib.gnca_profile_enable(1)
        nprof = min(steps, 5)
        for i in range(nprof):
            flush.fill_(1.0)
            one_rollout(x0_dev, scheds[warmup + i])
        torch.cuda.synchronize()
        lib.gnca_profile_enable(0)
        per_kernel = {}
        for kid, kname in ((0, "k_update"), (1, "k_apply"), (2, "k_rep_fwd"), (3, "k_rep_wgrad"), (6, "k_rep_bwd")):
            ms, n = ctypes.c_double(0), ctypes.c_ulonglong(0)
            lib.gnca_profile_read(kid, ctypes.byref(ms), ctypes.byref(n))
            if n.value:
                per_kernel[kname] = (ms.value, n.value)
        kernels_ms = {k: {"ms_per_step": v[0] / nprof, "launches_per_step": v[1] / nprof} for k, v in per_kernel.items()}
        if per_kernel:
            kname = max(per_kernel, key=lambda k: per_kernel[k][0])
            ms, n = per_kernel[kname]
            # algorithmic (dense) flops / launch.  Training step: the credited 3 x forward flops split evenly over its three
            # kernels (forward 17.9 k; data gradients gh, gy, perception^T, message^T 17.9 k; weight gradients 17.8 k per
            # cell-update -- the hidden-layer recompute inside k_rep_bwd is overhead and not credited)
            kflop = cfg["flop"] / 3.0 if cfg["train"] else cfg["flop"]
            flops_per_launch = kflop * updates * nprof / n
            avg_s = ms * 1e-3 / n
            clocks = sampler.result()
            mhz = clocks["sm_max_mhz"] or 1965
            peak = FP32_LANES * mhz * 1e6 / 1e12
            achieved = flops_per_launch / avg_s / 1e12
            roof = {"bound": "fp32_fma", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": NCU_DRAM_BYTES.get((cfg["name"][:2], kname)) if (cfg["name"][:2] != "c5" or B == 16) else None,
                    "fma_pipe_pct_ncu": NCU_FMA_PIPE_PCT.get((cfg["name"][:2], kname)), "avg_launch_ms": avg_s * 1e3, "launches_per_step": n / nprof,
                    "peak_source": f"derived: 148 SM x 128 FMA lanes x 2 x {mhz} MHz (MEASURED_PEAKS.json has no fp32 entry; "
                                   "hbm_gbs 6547.8 measured is far from binding: see hbm_view)",
                    "note": "achieved = DENSE algorithmic flops (every cell counted) / measured kernel time; the kernel "
                            "skips cells whose fire*alive mask is 0, so frac is a dense-equivalent figure",
                    # resident kernel: x_0 in + x_T out per rollout; streaming step kernels: state in + out per CA step
                    "hbm_view": {"algorithmic_bytes_per_step": int(2 * x0_host.numel() * 4 * (T if kname in ("k_update", "k_apply") else 1)),
                                 "achieved_GBps": 2 * x0_host.numel() * 4 * (T if kname in ("k_update", "k_apply") else 1) * nprof / (ms * 1e-3) / 1e9,
                                 "peak_GBps": 6547.8},
                    "share_of_step": ms / nprof / (dev_ms_max / steps), "kernels": kernels_ms}

    return {"value": value, "ms_per_step": dev_ms_max / steps, "e2e": e2e, "launches": int(launches), "roofline": roof,
            "clocks": sampler.result(), "data": data, "dims": (C_, H, W, B, T)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--rollout-impl", default="auto", choices=["auto", "streaming", "resident", "banded"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-extra", action="store_true", help="skip the fwd+bwd (c3) measurement of the default run")
    ap.add_argument("--T", type=int, default=0, help="development: override the number of CA steps per rollout")
    ap.add_argument("--B", type=int, default=0, help="development: override the per-GPU batch")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = workload_cfg(args.workload)
    if args.T > 0:
        cfg["T"] = args.T
    if args.B > 0:
        cfg["B"] = args.B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, cfg, rank)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    res = run_workload(cfg, args, world, rank, local_rank, dev, args.steps, args.warmup)
    C_, H, W, B, T = res["dims"]
    value, e2e, launches, roof, data = res["value"], res["e2e"], res["launches"], res["roofline"], res["data"]

    # the other half of the metric ("fwd, fwd+bwd"): the training hot path (fwd with BPTT records, loss, resident
    # backward, batched weight gradients, all-reduce, normalise + Adam) on BASELINE configs[2]'s shape, short regime
    train = None
    if args.workload == "c2" and not args.no_train_extra:
        cfg3 = workload_cfg("c3")
        r3 = run_workload(cfg3, args, world, rank, local_rank, dev, max(3, min(args.steps, 10)), 3, with_roofline=True)
        train = {"metric": "graph-NCA cell-updates/s (fwd+bwd)", "value": r3["value"], "unit": "cell-updates/s",
                 "ms_per_step": r3["ms_per_step"], "e2e": r3["e2e"], "gpu_launches": r3["launches"],
                 "config": {"workload": cfg3["name"], "B_per_gpu": cfg3["B"], "T": cfg3["T"], "fire_rate": cfg3["fire_rate"],
                            "message_every": cfg3["message_every"]},
                 "kernels": (r3["roofline"] or {}).get("kernels")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not cfg["train"] and not cfg.get("classic"):
        cpu = time_cpu_port(cfg)

    if rank == 0:
        line = {"metric": "graph-NCA cell-updates/s (%s)" % ("fwd+bwd" if cfg["train"] else "fwd"), "value": value,
                "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": data,
                "config": {"workload": cfg["name"], "B_per_gpu": B, "T": T, "grid": [H, W], "channels": C_,
                           "hidden": cfg["hidden"], "fire_rate": cfg["fire_rate"], "offsets_per_step": 8,
                           "graph_shift": "torus", "fire_rng": "in-kernel philox", "rollout_impl": args.rollout_impl,
                           "l2": "flushed between timed iterations (256 MiB fill)", "parallelism": f"dp{world} batch-sharded, no collective"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": res["clocks"], "roofline": roof,
                "cpu_baseline": cpu, "fwd_bwd": train}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
