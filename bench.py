#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched 3D bin-packing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload target|cfg1|cfg2|cfg3|cfg4]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the CPU restatement of the reference path, all host threads

One "step" = one pass of the hot path over every bin of this rank: action in -> placement check,
heightmap update, reward/done/info, auto-reset, next item, float32 observation, float32 ground-truth
mask, and the next uniformly-random valid action (the policy stand-in, fused into the same launch).

  value      env-steps/s with all inputs resident in HBM (CUDA events around K steps, max over ranks)
  e2e        the same through the reference-facing host API: each step the actions leave the device
             (VecPyTorch.step_async's .cpu(), acktr/envs.py:182-187), come back through
             bpp_step_host() from pinned host memory, and reward/done/counter/ratio return to the host
             (acktr/envs.py:189-193); obs and mask stay on the device as in the reference.
  roofline   algorithmic bytes/step ((24+4R)*L*W+25 per env-step) / measured time,
             against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  the oracle (C restatement of the reference path, OpenMP over bins) on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "online-3d-bpp-drl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (bins per GPU, container, rotation, generator, description)
    "target": (1 << 19, (10, 10, 10), False, "rs", "2^19 parallel 10x10x10 bins per GPU, rs generator (north_star target config)"),
    "target_rot": (1 << 19, (10, 10, 10), True, "rs", "2^19 parallel 10x10x10 bins per GPU, rs generator, rotation enabled"),
    "cfg1": (4096, (10, 10, 10), False, "rs", "BASELINE configs[1]: 4096 bins, 10x10x10, rs"),
    "cfg2": (65536, (10, 10, 10), False, "cut2_replay", "BASELINE configs[2]: 65536 bins, 10x10x10, cut_2.pt replay"),
    "cfg3": (262144, (20, 20, 20), False, "rs", "BASELINE configs[3]: 262144 bins, 20x20x20, rs"),
    "cut1": (65536, (10, 10, 10), False, "cut1", "65536 bins, 10x10x10, cut1 generator on device"),
    "cut2": (65536, (10, 10, 10), False, "cut2", "65536 bins, 10x10x10, cut2 generator on device"),
    "cfg4": (65536, (10, 10, 10), True, "cut2_replay", "BASELINE configs[4]: 65536 bins per GPU (524288 over 8), rotation, cut_2 replay"),
}


def algorithmic_bytes(size, rot):
    A = size[0] * size[1]
    return (24 + 4 * (2 if rot else 1)) * A + 25


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_per_step():
    """dram bytes per launch from the committed ncu capture, if one was recorded for this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def gen_kwargs(gen):
    if gen == "cut2_replay":
        d = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
        return dict(trajectories=(d["cut_2"], d["cut_2_len"]), first_traj=1)
    return dict(data_type=gen)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.out,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, mx, reasons = [], [], set()
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_rate(size, rot, gen, seed, budget_s, n_threads=0, bins=65536, steps=None, warmup=1):
    """Times the oracle (oracle/bpp_oracle.c, OpenMP) on a bounded sample of the workload."""
    import oracle as orc
    if n_threads <= 0:      # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core it is allowed to run on
        try:
            n_threads = len(os.sched_getaffinity(0))
        except AttributeError:
            n_threads = os.cpu_count() or 1
    kw = gen_kwargs(gen)
    if "trajectories" in kw:
        tr = kw.pop("trajectories")
        okw = dict(gen="replay", traj=tr[0], traj_len=tr[1], first_traj=kw.get("first_traj", 1))
    else:
        okw = dict(gen=gen)
    env = orc.OracleEnv(bins, size, rot=rot, seed=seed, n_threads=n_threads, **okw)
    env.reset()
    t_warm = time.perf_counter()
    n_warm = 0
    while n_warm < warmup or time.perf_counter() - t_warm < 1.5:     # OpenMP pool spin-up, page faults, clocks
        env.step(env.sample.copy())
        n_warm += 1
    t0 = time.perf_counter()
    done_steps = 0
    while True:                      # fixed step count (reference arm) or a time budget (cpu_baseline of our arm)
        env.step(env.sample.copy())
        done_steps += 1
        if steps is not None:
            if done_steps >= steps:
                break
        elif time.perf_counter() - t0 >= budget_s and done_steps >= 3:
            break
    dt = time.perf_counter() - t0
    steps = done_steps
    cores = n_threads
    env.close()
    return {"value": bins * steps / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": "%d bins x %d steps of the same config (%s), oracle/bpp_oracle.c with OpenMP over bins, %.1f s"
                      % (bins, steps, gen, dt)}, steps, dt


def aux_small_config(b200bpp, _cabi, dev, seed, steps=300, warmup=20):
    """BASELINE configs[1] (4096 bins, 10x10x10, rs) beside the headline workload: launch-latency bound."""
    import torch
    bins, size, rot, gen, desc = WORKLOADS["cfg1"]
    env = b200bpp.BinPackingVecEnv(bins, size, enable_rotation=rot, seed=seed, device=dev, want_sample=True, host_infos=False)
    lib = _cabi.lib()
    stream = torch.cuda.current_stream(dev)
    P = lambda t: C.c_void_p(t.data_ptr())           # noqa: E731
    bufs = [env.sample_t, torch.zeros_like(env.sample_t)]
    env.reset()

    def step(i):
        _cabi.check(lib.bpp_step(env._h, P(bufs[i & 1]), P(env.obs), P(env.location_masks), P(env.reward_t), P(env.done_t),
                                 P(env.counter_t), P(env.ratio_t), P(bufs[(i + 1) & 1]), C.c_void_p(stream.cuda_stream)))
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(warmup, warmup + steps):
        step(i)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    env.close()
    return {"workload": desc, "ms_per_step": ms, "value": bins / (ms * 1e-3), "unit": "env-steps/s", "steps": steps,
            "note": "one kernel launch per step; at 4096 bins the step is launch-latency bound"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    bins, size, rot, gen, desc = WORKLOADS[args.workload]
    sample_bins = min(bins, 65536)
    base, steps, dt = cpu_oracle_rate(size, rot, gen, args.seed, budget_s=0, bins=sample_bins,
                                      steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": base["value"], "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": desc, "bins_per_step_sample": sample_bins, "container": list(size), "rotation": rot,
                   "generator": gen, "note": "reference is pure Python and cannot travel to the GPU box; this is its C restatement "
                                            "(oracle/), every host thread, host memory only"},
        "cpu_baseline": dict(base),
        "e2e": {"value": base["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--bins-per-gpu", type=int, default=None)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import b200bpp
    from b200bpp import _cabi

    bins, size, rot, gen, desc = WORKLOADS[args.workload]
    if args.bins_per_gpu:
        bins = args.bins_per_gpu
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    env = b200bpp.BinPackingVecEnv(bins, size, enable_rotation=rot, seed=args.seed, device=dev,
                                   global_offset=rank * bins, global_n=world * bins, want_sample=True,
                                   host_infos=False, **gen_kwargs(gen))
    lib = _cabi.lib()
    N = bins
    sample_a = env.sample_t
    sample_b = torch.zeros_like(sample_a)
    stream = torch.cuda.current_stream(dev)
    sptr = lambda: C.c_void_p(stream.cuda_stream)   # noqa: E731
    P = lambda t: C.c_void_p(t.data_ptr())           # noqa: E731

    def step_device(src, dst):
        _cabi.check(lib.bpp_step(env._h, P(src), P(env.obs), P(env.location_masks), P(env.reward_t), P(env.done_t),
                                 P(env.counter_t), P(env.ratio_t), P(dst), sptr()))

    env.reset()
    bufs = [sample_a, sample_b]
    for i in range(args.warmup):
        step_device(bufs[i & 1], bufs[(i + 1) & 1])
    cur = args.warmup & 1

    def barrier():
        torch.cuda.synchronize(dev)
        if dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing --------------------------------------------------------------------
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = lib.bpp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_device(bufs[(cur + i) & 1], bufs[(cur + i + 1) & 1])
    e1.record(stream)
    barrier()
    launches = lib.bpp_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    cur = (cur + args.steps) & 1

    # ---- end to end through the host API -------------------------------------------------------------
    e2e_steps = args.e2e_steps or max(10, min(args.steps, 100))
    h_act = torch.empty((N,), dtype=torch.int64).pin_memory()
    h_rew = torch.empty((N,), dtype=torch.float32).pin_memory()
    h_done = torch.empty((N,), dtype=torch.uint8).pin_memory()
    h_cnt = torch.empty((N,), dtype=torch.int32).pin_memory()
    h_rat = torch.empty((N,), dtype=torch.float32).pin_memory()

    h_act2 = torch.empty((N,), dtype=torch.int64).pin_memory()
    hbufs = [h_act, h_act2]

    def step_host(src, dst):
        # actions come from pinned host memory; the env writes the next sampled action (the policy stand-in's
        # output) straight into the other pinned buffer, i.e. the policy output also crosses PCIe every step
        _cabi.check(lib.bpp_step_host(env._h, P(src), P(env.obs), P(env.location_masks), P(h_rew), P(h_done),
                                      P(h_cnt), P(h_rat), P(dst), sptr()))

    hbufs[cur & 1].copy_(bufs[cur & 1])
    torch.cuda.synchronize(dev)
    bufs = hbufs
    for i in range(3):
        step_host(bufs[(cur + i) & 1], bufs[(cur + i + 1) & 1])
    cur = (cur + 3) & 1
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record(stream)
    for i in range(e2e_steps):
        step_host(bufs[(cur + i) & 1], bufs[(cur + i + 1) & 1])
    f1.record(stream)
    barrier()
    e2e_wall_ms = 1e3 * (time.perf_counter() - t0)
    e2e_ms = max(f0.elapsed_time(f1), 0.0)
    done_frac = float(h_done.float().mean())

    # ---- max over ranks; the single metrics collective ------------------------------------------------
    t = torch.tensor([ms, e2e_ms, e2e_wall_ms], dtype=torch.float64, device=dev)
    m = env.metrics()
    met = torch.tensor([m["episodes"], m["sum_counter"], m["sum_vol"], m["steps"]], dtype=torch.int64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(met, op=dist.ReduceOp.SUM)          # the one NCCL collective of the path
    ms, e2e_ms, e2e_wall_ms = (float(v) for v in t.tolist())
    met = [int(v) for v in met.tolist()]

    if rank == 0:
        total_bins = N * world
        value = total_bins * args.steps / (ms * 1e-3)
        e2e_value = total_bins * e2e_steps / (max(e2e_ms, e2e_wall_ms) * 1e-3)
        b_alg = algorithmic_bytes(size, rot)
        peak, peak_src = hbm_peak()
        achieved = N * b_alg / (ms * 1e-3 / args.steps) / 1e9          # per GPU, per launch
        tr = traffic_per_step().get(args.workload)
        LWH = size[0] * size[1] * size[2]
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 state / int32 arithmetic / f32 outputs", "data": "synthetic",
            "config": {"workload": desc, "bins_per_gpu": N, "total_bins": total_bins, "container": list(size),
                       "rotation": rot, "generator": gen, "mask_variant": "v2 (acktr.utils.check_box)",
                       "policy": "uniform-random valid action, sampled in the step kernel",
                       "build": lib.bpp_build_info().decode(),
                       "parallelism": "bins sharded %d-way, no data-path collective" % world,
                       "l2": "per-step working set %.0f MB per GPU vs 126 MB L2 (inputs larger than L2, no flush)"
                             % (N * (b_alg) / 1e6)},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "steps": e2e_steps,
                    "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": (8 + 4 + 1 + 4 + 4) * N,
                    "ms_per_step": max(e2e_ms, e2e_wall_ms) / e2e_steps,
                    "api": "bpp_step_host (the call under BinPackingVecEnv.step): int64 actions read from pinned host "
                           "memory, reward/done/counter/ratio and the next sampled action written to pinned host memory "
                           "(zero-copy over PCIe inside the kernel), stream synchronised every step; obs+mask stay on "
                           "the device as in acktr/envs.py:191"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": tr, "peak_source": peak_src, "algorithmic_bytes_per_env_step": b_alg,
                         "kernel": "bpp_step_kernel (one launch per step)"},
            "clocks": clk,
            "episodes": {"finished": met[0], "mean_items": met[1] / max(met[0], 1),
                         "mean_utilisation": met[2] / max(met[0], 1) / LWH, "done_frac_last_step": done_frac},
        }
        if world == 1 and args.workload == "target":
            line["aux"] = {"configs[1]": aux_small_config(b200bpp, _cabi, dev, args.seed)}
        if world == 1 and not args.no_cpu_baseline:
            base, _, _ = cpu_oracle_rate(size, rot, gen, args.seed, budget_s=12.0)
            line["cpu_baseline"] = base
        print(json.dumps(line))
    env.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
