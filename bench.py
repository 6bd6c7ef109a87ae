#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched 3D bin-packing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload target|cfg1|cfg2|cfg3|cfg4]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the CPU restatement of the reference path, all host threads

One "step" = one pass of the hot path over every bin of this rank: action in -> placement check,
heightmap update, reward/done/info, auto-reset, next item, float32 observation, float32 ground-truth
mask, and the next uniformly-random valid action (the policy stand-in, fused into the same launch).

  value      env-steps/s with all inputs resident in HBM (CUDA events around K steps, max over ranks)
  e2e        the same through the C ABI with HOST buffers (bpp_step_packed_host_i32): every step the int32 actions are read
             from pinned host memory and one 16-byte result record + one done byte per bin are written to pinned host
             memory (what VecPyTorch.step_async / step_wait move, acktr/envs.py:182-193); obs and mask stay on the
             device as in the reference.  e2e_py is the public Python call, BinPackingVecEnv.step(), timed the same way.
  roofline   algorithmic bytes/step ((24+4R)*L*W+25 per env-step, reference dtypes) / measured time,
             against MEASURED_PEAKS.json's HBM copy bandwidth; `dram` is the same with the bytes the kernel really
             moves (u8 state; profiles/traffic.json, from ncu).
  aux        every other BASELINE.json config, each with its own roofline; at every N configs[4] strong-scaled
             (2^19 bins over all ranks, rotation, cut_2 replay, one NCCL reduce of the episode metrics).
  cpu_baseline / --impl reference   the UNMODIFIED reference on this box's host cores (oracle/_ref, vendored byte for
             byte by oracle/make_ref.py): make_vec_envs -> ShmemVecEnv(fork) + the host mask loop of main.py:163-169,
             SubprocVecEnv beside it; the C/OpenMP port of the path (oracle/) as a second, labelled figure.
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
CFG5_TOTAL = 1 << 19          # BASELINE configs[4]: 524288 bins over all GPUs of the box (strong scaling)


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


def host_threads():
    """Threads the CPU arm may really use: the affinity mask capped by the cgroup CPU quota."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_bench
    return ref_bench.host_cores()


def cpu_port_rate(size, rot, gen, seed, budget_s, bins, n_threads=0, steps=None, warmup=1):
    """The C/OpenMP port of the reference path (oracle/bpp_oracle.c) on a bounded number of steps of the workload."""
    import oracle as orc
    how = None
    if n_threads <= 0:      # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core it is allowed to run on
        n_threads, how = host_threads()
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
    while n_warm < warmup or time.perf_counter() - t_warm < 1.0:     # OpenMP pool spin-up, page faults, clocks
        env.step(env.sample.copy())
        n_warm += 1
    t0 = time.perf_counter()
    done_steps = 0
    while True:
        env.step(env.sample.copy())
        done_steps += 1
        if steps is not None:
            if done_steps >= steps:
                break
        elif time.perf_counter() - t0 >= budget_s and done_steps >= 3:
            break
    dt = time.perf_counter() - t0
    env.close()
    return {"value": bins * done_steps / dt, "unit": "env-steps/s", "cores": n_threads, "cores_detail": how, "kind": "port",
            "sample": "%d bins x %d steps of the same config (%s), oracle/bpp_oracle.c with OpenMP over bins, %.1f s"
                      % (bins, done_steps, gen, dt)}


def workload_config(name, world, bins_per_gpu=None):
    """The `config` of the JSON line -- the same dict in both arms (the reference arm times a bounded sample of it)."""
    bins, size, rot, gen, desc = WORKLOADS[name]
    bins = bins_per_gpu or bins
    b_alg = algorithmic_bytes(size, rot)
    return {"workload": desc, "bins_per_gpu": bins, "total_bins": bins * world, "container": list(size), "rotation": rot,
            "generator": gen, "mask_variant": "v2 (acktr.utils.check_box)",
            "policy": "uniform-random valid action per bin",
            "parallelism": "bins sharded %d-way, no data-path collective" % world,
            "l2": "per-step working set %.0f MB per GPU vs 126 MB L2 (inputs larger than L2, no flush)" % (bins * b_alg / 1e6)}


def run_reference(args, rank, world):
    """The reference arm: the UNMODIFIED reference's own CPU path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    bins, size, rot, gen, desc = WORKLOADS[args.workload]
    ref_gen = {"cut2_replay": "cut2"}.get(gen, gen)        # make_vec_envs has no dataset replay (test=False, acktr/envs.py:39)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_bench
    extra = {}
    try:
        base = ref_bench.time_shmem(size, rot, ref_gen, seconds=args.ref_seconds, seed=args.seed, steps=args.steps,
                                    warmup_steps=args.warmup)
        dt = base["seconds"]
        try:
            extra["subproc_step_only"] = ref_bench.time_subproc(size, rot, ref_gen, seconds=3.0, seed=args.seed)
            extra["subproc_step_and_mask"] = ref_bench.time_subproc(size, rot, ref_gen, seconds=3.0, seed=args.seed, with_mask=True)
        except Exception as e:      # noqa: BLE001
            extra["subproc_error"] = repr(e)
    except RuntimeError as e:       # neither /root/reference nor oracle/_ref: fall back to the port, and say so
        base = cpu_port_rate(size, rot, gen, args.seed, budget_s=0, bins=min(bins, 65536), steps=args.steps, warmup=args.warmup)
        base["note"] = "reference tree absent (%s): C/OpenMP port timed instead" % e
        dt = None
    if not args.no_port:
        # second, clearly labelled figure: the C/OpenMP port on the full-size config
        extra["port_openmp_full_config"] = cpu_port_rate(size, rot, gen, args.seed, budget_s=6.0, bins=bins)
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": base["value"], "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps if dt else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32 / float64 (numpy)", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus, args.bins_per_gpu),
        "cpu_baseline": dict(base),
        "e2e": {"value": base["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line.update(extra)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
class Runner(object):
    """One env on this rank driven through the C ABI with the in-kernel sampled action as the policy stand-in."""

    def __init__(self, b200bpp, _cabi, name, dev, seed, rank, world, bins=None):
        import torch
        self.torch, self.lib, self._cabi = torch, _cabi.lib(), _cabi
        b, self.size, self.rot, self.gen, self.desc = WORKLOADS[name]
        self.name = name
        self.N = bins or b
        self.dev = dev
        self.env = b200bpp.BinPackingVecEnv(self.N, self.size, enable_rotation=self.rot, seed=seed, device=dev,
                                            global_offset=rank * self.N, global_n=world * self.N, want_sample=True,
                                            host_infos=False, **gen_kwargs(self.gen))
        self.stream = torch.cuda.current_stream(dev)
        self.bufs = [self.env.sample_t, torch.zeros_like(self.env.sample_t)]
        self.k = 0
        self.env.reset()

    def P(self, t):
        return C.c_void_p(t.data_ptr())

    def step(self):
        e, src, dst = self.env, self.bufs[self.k & 1], self.bufs[(self.k + 1) & 1]
        self._cabi.check(self.lib.bpp_step(e._h, self.P(src), self.P(e.obs), self.P(e.location_masks), self.P(e.reward_t),
                                           self.P(e.done_t), self.P(e.counter_t), self.P(e.ratio_t), self.P(dst),
                                           C.c_void_p(self.stream.cuda_stream)))
        self.k += 1

    def time_device(self, steps, warmup, barrier=None):
        torch = self.torch
        for _ in range(warmup):
            self.step()
        (barrier or (lambda: torch.cuda.synchronize(self.dev)))()
        launches0 = self.lib.bpp_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            self.step()
        e1.record(self.stream)
        (barrier or (lambda: torch.cuda.synchronize(self.dev)))()
        return e0.elapsed_time(e1) / steps, self.lib.bpp_launch_count() - launches0

    def roofline(self, ms, traffic=None, kernel="bpp_step_kernel (one launch per step)"):
        b_alg = algorithmic_bytes(self.size, self.rot)
        peak, peak_src = hbm_peak()
        achieved = self.N * b_alg / (ms * 1e-3) / 1e9
        out = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
               "peak_source": peak_src, "algorithmic_bytes_per_env_step": b_alg, "kernel": kernel}
        if traffic:     # the bytes the kernel really moves (u8 state): ncu dram__bytes_read + write per launch of the
            # workload's standard size, scaled to this env's bin count
            traffic = traffic * self.N / float(WORKLOADS[self.name][0])
            out["traffic"] = traffic
            real = traffic / (ms * 1e-3) / 1e9
            out["dram"] = {"achieved": real, "frac": real / peak, "bytes_per_env_step": traffic / self.N,
                           "note": "real DRAM traffic of the launch (ncu) / CUDA-event time; u8 heightmaps move fewer bytes than "
                                   "the reference-dtype accounting of `frac`"}
        return out

    def close(self):
        self.env.close()


def time_e2e(r, steps, warmup=3):
    """(wall ms per step, done fraction) through the C ABI with NUMA-local pinned HOST buffers, stream synchronised every
    step.  Every step the kernel reads its int32 actions from pinned host memory -- the next_action column of the records
    the previous step wrote there: the policy stand-in's output, handed back without a host pass -- and writes one 16-byte
    record + one done byte per bin to pinned host memory."""
    import torch
    from b200bpp.vec_env import _HostBuffer
    lib, N, dev = r.lib, r.N, r.dev
    hb = [_HostBuffer(lib, dev.index, 16 * N), _HostBuffer(lib, dev.index, 16 * N), _HostBuffer(lib, dev.index, N)]
    rec = [h.view(np.int32, (N, 4)) for h in hb[:2]]
    rec[0][:, 3] = r.bufs[r.k & 1].cpu().numpy().astype(np.int32)      # the env's pending sampled actions
    torch.cuda.synchronize(dev)
    k = 0

    def step():
        nonlocal k
        src, dst = hb[k & 1], hb[(k + 1) & 1]
        r._cabi.check(lib.bpp_step_packed_host_i32(r.env._h, C.c_void_p(src.ptr + 12), 4, r.P(r.env.obs), r.P(r.env.location_masks),
                                                   C.c_void_p(dst.ptr), C.c_void_p(hb[2].ptr), None, C.c_void_p(r.stream.cuda_stream)))
        k += 1
    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize(dev)
    wall = 1e3 * (time.perf_counter() - t0) / steps
    done_frac = float(hb[2].view(np.uint8, (N,)).mean())
    # leave the device-side action buffer in step with the env
    r.bufs[r.k & 1].copy_(torch.from_numpy(rec[k & 1][:, 3].astype(np.int64)))
    torch.cuda.synchronize(dev)
    for h in hb:
        h.free()
    return wall, done_frac


def time_e2e_py(r, steps, warmup=3):
    """ms per call of the public Python API: obs, reward, done, infos = env.step(actions), actions = the device tensor a
    GPU policy would hand over; reward / done / infos arrive in pinned host memory."""
    import torch
    e = r.env
    a = r.bufs[r.k & 1]
    if a is not e.sample_t:
        e.sample_t.copy_(a)
    torch.cuda.synchronize(r.dev)
    for _ in range(warmup):
        e.step(e.sample_t.view(-1, 1))
    t0 = time.perf_counter()
    for _ in range(steps):
        obs, reward, done, infos = e.step(e.sample_t.view(-1, 1))
    wall = 1e3 * (time.perf_counter() - t0) / steps
    r.bufs, r.k = [e.sample_t, r.bufs[1] if r.bufs[0] is e.sample_t else r.bufs[0]], 0
    return wall


def aux_line(r, ms, extra=None):
    out = {"workload": r.desc, "bins_per_gpu": r.N, "ms_per_step": ms, "value": r.N / (ms * 1e-3), "unit": "env-steps/s",
           "roofline": r.roofline(ms, traffic_per_step().get(r.name))}
    out.update(extra or {})
    return out


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
    ap.add_argument("--no-aux", action="store_true")
    ap.add_argument("--no-port", action="store_true", help="reference arm: skip the C/OpenMP port figure")
    ap.add_argument("--ref-seconds", type=float, default=15.0, help="reference arm: seconds of reference stepping")
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

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _cabi.lib().bpp_bind_thread_to_device(local_rank)      # launch thread next to its GPU (NUMA); harmless if it fails
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    # ---- headline: device-resident --------------------------------------------------------------------------------
    r = Runner(b200bpp, _cabi, args.workload, dev, args.seed, rank, world, args.bins_per_gpu)
    clocks = ClockSampler(local_rank)
    for _ in range(args.warmup):
        r.step()
    barrier()
    if rank == 0:
        clocks.start()
    ms, launches = r.time_device(args.steps, 0, barrier)
    clk = clocks.stop() if rank == 0 else None

    # ---- end to end: C ABI with host buffers, then the public Python call ----------------------------------------------
    e2e_steps = args.e2e_steps or max(10, min(args.steps, 100))
    barrier()
    e2e_ms, done_frac = time_e2e(r, e2e_steps)
    barrier()
    e2e_py_ms = time_e2e_py(r, e2e_steps)
    barrier()
    ms, e2e_ms, e2e_py_ms = max_over_ranks([ms, e2e_ms, e2e_py_ms])

    # ---- the one collective of the path: episode metrics ------------------------------------------------------------
    m = r.env.metrics()
    met = torch.tensor([m["episodes"], m["sum_counter"], m["sum_vol"], m["steps"]], dtype=torch.int64, device=dev)
    if dist:
        dist.all_reduce(met, op=dist.ReduceOp.SUM)
    met = [int(v) for v in met.tolist()]
    N, size = r.N, r.size
    LWH = size[0] * size[1] * size[2]

    # ---- the other BASELINE configs -------------------------------------------------------------------------------------
    aux = {}
    if not args.no_aux and args.workload == "target":
        # configs[4] strong-scaled: 2^19 bins over all ranks, rotation, cut_2 replay, one NCCL reduce of the metrics
        r5 = Runner(b200bpp, _cabi, "cfg4", dev, args.seed, rank, world, bins=CFG5_TOTAL // world)
        ms5, _ = r5.time_device(200, 20, barrier)
        m5 = r5.env.metrics()
        t5 = torch.tensor([m5["episodes"], m5["sum_counter"], m5["sum_vol"]], dtype=torch.int64, device=dev)
        if dist:
            dist.all_reduce(t5, op=dist.ReduceOp.SUM)
        (ms5,) = max_over_ranks([ms5])
        ep5 = [int(v) for v in t5.tolist()]
        aux["configs[4]"] = aux_line(r5, ms5, {
            "workload": "BASELINE configs[4]: 524288 bins sharded over %d GPU(s), 10x10x10, rotation, cut_2 replay" % world,
            "total_bins": CFG5_TOTAL, "value": CFG5_TOTAL / (ms5 * 1e-3), "scaling": "strong",
            "episodes": {"finished": ep5[0], "mean_items": ep5[1] / max(ep5[0], 1), "mean_utilisation": ep5[2] / max(ep5[0], 1) / 1000.0},
            "collective": "one NCCL all-reduce (sum) of 3 int64 episode accumulators after the timed region"})
        r5.close()
        if world == 1:
            for key, name, steps in (("configs[1]", "cfg1", 400), ("configs[2]", "cfg2", 300), ("configs[3]", "cfg3", 100),
                                     ("target_rot", "target_rot", 100), ("cut1", "cut1", 100), ("cut2", "cut2", 100)):
                ra = Runner(b200bpp, _cabi, name, dev, args.seed, 0, 1)
                msa, _ = ra.time_device(steps, 20)
                extra = {}
                if name == "cfg1":
                    # small envs are launch-bound: K env steps per launch (bpp_step_packed with n_steps = K)
                    extra["multi_step"] = time_multi_step(ra, K=16, launches=40)
                    extra["cuda_graph"] = time_graph(ra, K=16, replays=40)
                    extra["note"] = ("one launch per step is launch-latency bound at 4096 bins; multi_step runs 16 env steps per "
                                     "launch, cuda_graph replays a captured graph of 16 single-step launches")
                aux[key] = aux_line(ra, msa, extra)
                ra.close()

    if rank == 0:
        total_bins = N * world
        value = total_bins / (ms * 1e-3)
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 state / int32 arithmetic / f32 outputs", "data": "synthetic",
            "config": workload_config(args.workload, world, args.bins_per_gpu),
            "build": r.lib.bpp_build_info().decode(),
            "e2e": {"value": total_bins / (e2e_ms * 1e-3), "unit": "env-steps/s", "steps": e2e_steps,
                    "h2d_bytes_per_step": 4 * N, "d2h_bytes_per_step": 17 * N, "ms_per_step": e2e_ms,
                    "api": "bpp_step_packed_host_i32 (bpp_step_packed_host, the call under BinPackingVecEnv.step, with int32 "
                           "actions): actions read from pinned, NUMA-local host memory (the next_action column of the previous "
                           "step's records: 4 bytes per bin are consumed, the strided read touches the whole 16-byte "
                           "records), one 16-byte result record (reward, placed volume, counter, next action) + one done "
                           "byte per bin written to pinned host memory, zero-copy over PCIe inside the kernel, stream "
                           "synchronised every step; obs+mask stay on the device as in acktr/envs.py:191; wall clock, max over ranks"},
            "e2e_py": {"value": total_bins / (e2e_py_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": e2e_py_ms,
                       "api": "obs, reward, done, infos = BinPackingVecEnv.step(actions) with the actions a GPU policy hands "
                              "over (device int64[N,1]); reward float32[N,1], done np.bool_[N] and the lazy infos are views "
                              "of the pinned records; wall clock, max over ranks"},
            "gpu_launches": int(launches),
            "roofline": r.roofline(ms, traffic_per_step().get(args.workload)),
            "clocks": clk,
            "episodes": {"finished": met[0], "mean_items": met[1] / max(met[0], 1),
                         "mean_utilisation": met[2] / max(met[0], 1) / LWH, "done_frac_last_step": done_frac},
        }
        if aux:
            line["aux"] = aux
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = reference_subprocess(args)
        print(json.dumps(line))
    r.close()
    if dist:
        dist.destroy_process_group()


def time_multi_step(r, K, launches):
    """env-steps/s with K env steps per launch (bpp_step_packed, n_steps = K): policy-free rollouts on small envs."""
    import torch
    e, dev = r.env, r.dev
    res = torch.zeros((r.N, 4), dtype=torch.int32, device=dev)

    def launch():
        r._cabi.check(r.lib.bpp_step_packed(e._h, r.P(e.sample_t), r.P(e.obs), r.P(e.location_masks), r.P(res), r.P(e.done_t),
                                            r.P(e.sample_t), K, C.c_void_p(r.stream.cuda_stream)))
    a = r.bufs[r.k & 1]
    if a is not e.sample_t:
        e.sample_t.copy_(a)
    for _ in range(5):
        launch()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(r.stream)
    for _ in range(launches):
        launch()
    e1.record(r.stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / (launches * K)
    r.bufs, r.k = [e.sample_t, r.bufs[1] if r.bufs[0] is e.sample_t else r.bufs[0]], 0
    return {"env_steps_per_launch": K, "ms_per_env_step": ms, "value": r.N / (ms * 1e-3), "unit": "env-steps/s"}


def time_graph(r, K, replays):
    """env-steps/s with K single-step launches captured into one CUDA graph (in-place action buffer) and replayed."""
    import torch
    e, dev = r.env, r.dev
    a = r.bufs[r.k & 1]
    if a is not e.sample_t:
        e.sample_t.copy_(a)
    torch.cuda.synchronize(dev)
    s = torch.cuda.Stream(device=dev)
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.stream(s):
            e.step_device(e.sample_t)
            torch.cuda.synchronize(dev)
            with torch.cuda.graph(g, stream=s):
                for _ in range(K):
                    e.step_device(e.sample_t)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(replays):
                g.replay()
            e1.record(s)
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / (replays * K)
    except Exception as ex:      # noqa: BLE001
        return {"error": repr(ex)}
    r.bufs, r.k = [e.sample_t, r.bufs[1] if r.bufs[0] is e.sample_t else r.bufs[0]], 0
    return {"env_steps_per_graph": K, "ms_per_env_step": ms, "value": r.N / (ms * 1e-3), "unit": "env-steps/s"}


def reference_subprocess(args):
    """cpu_baseline of our arm: the reference arm in a fresh process (fork-based workers, no CUDA context around)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "10",
           "--warmup", "2", "--ref-seconds", "12", "--seed", str(args.seed)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=240,
                             env={k: v for k, v in os.environ.items() if k != "OMP_NUM_THREADS"})
        for ln in out.stdout.splitlines():
            if ln.startswith("{"):
                d = json.loads(ln)
                base = d["cpu_baseline"]
                for k in ("subproc_step_only", "subproc_step_and_mask", "port_openmp_full_config"):
                    if k in d:
                        base[k] = d[k]
                return base
        return {"error": "no JSON line from the reference arm", "stderr": out.stderr[-400:]}
    except Exception as e:      # noqa: BLE001
        return {"error": repr(e)}


if __name__ == "__main__":
    main()
