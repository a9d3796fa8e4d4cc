#!/usr/bin/env python
"""Throughput benchmark of the retargeting hot path (hand-frames/s), one process per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference-side CPU path, same metric / config

Workload (BASELINE.json metric: "hand-frames/sec (21-kpt -> Allegro 16-DoF), batch 65536"): VectorOptimizer,
Allegro right hand, the shipped teleop config (scaling 1.6, huber 0.02, norm_delta 4e-3), one step = one batch
of 65 536 synthetic 21-keypoint frames PER GPU -> 16 joint angles each (weak scaling: frames shard across
ranks, no data-path collective).  Synthetic data as SURVEY.md section 8(d)(2): q* ~ U(limits), FK, wrist/tips
written at keypoints {0,4,8,12,16} divided by the scale (reachable), warm start q* + 0.05 N(0,1) clipped.

One JSON line on stdout (rank 0).  `value` = device-timed throughput with inputs resident in HBM; `e2e` = the
same batch through the host-buffer C-ABI call (pinned host memory in, pinned host memory out, copies inside
the timed region); `roofline` = algorithmic HBM bytes / measured launch time vs the measured copy peak;
`cpu_baseline` = the oracle's reference-faithful CPU path (scipy SLSQP at the reference's ftol) on this box's
host cores, on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

CONFIG_KEY = "teleop/allegro_hand_right"
FRAMES_PER_GPU = 65536
N_INPUT_SETS = 8  # rotating input batches: 8 x 24.9 MB > 126 MB L2
BYTES_PER_FRAME = 21 * 3 * 4 + 16 * 4 + 16 * 4  # keypoints in + warm start in + qpos out


# --------------------------------------------------------------------------------------- data
def make_batch(kin, opt_cfg, n, seed, centre=True):
    """Synthetic keypoint frames + warm starts (numpy, host)."""
    rng = np.random.RandomState(seed)
    lim = kin.joint_limits
    dof = kin.dof
    q = rng.uniform(lim[:, 0], lim[:, 1], size=(n, dof))
    init = np.clip(q + 0.05 * rng.randn(n, dof), lim[:, 0], lim[:, 1]).astype(np.float32)
    # batched float64 FK of the tips (data generation only)
    Rw = np.zeros((n, dof, 3, 3))
    pw = np.zeros((n, dof, 3))
    for i in range(dof):
        a = kin.joint_axis[i]
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        Rq = np.eye(3)[None] + np.sin(q[:, i])[:, None, None] * K[None] + (1 - np.cos(q[:, i]))[:, None, None] * (K @ K)[None]
        par = kin.joint_parent[i]
        if par >= 0:
            Rb = Rw[:, par] @ kin.joint_R[i]
            pb = np.einsum("bij,j->bi", Rw[:, par], kin.joint_p[i]) + pw[:, par]
        else:
            Rb = np.broadcast_to(kin.joint_R[i], (n, 3, 3))
            pb = np.broadcast_to(kin.joint_p[i], (n, 3))
        if kin.joint_type[i] == 0:
            Rw[:, i] = Rb @ Rq
            pw[:, i] = pb
        else:  # prismatic
            Rw[:, i] = Rb
            pw[:, i] = pb + np.einsum("bij,j->bi", Rb, a) * q[:, i:i + 1]
    kp = np.zeros((n, 21, 3), dtype=np.float32)
    names, human, scale = opt_cfg
    for name, h in zip(names, human):
        li = kin.link_index(name)
        par = kin.link_parent[li]
        if par >= 0:
            pos = np.einsum("bij,j->bi", Rw[:, par], kin.link_p[li]) + pw[:, par]
        else:
            pos = np.broadcast_to(kin.link_p[li], (n, 3))
        kp[:, h] = (pos / scale).astype(np.float32)
    if centre:  # make the origin (wrist) the zero of the keypoint frame, like a wrist-centred detector output
        kp -= kp[:, 0:1].copy()
    return kp, init


def workload(seq):
    opt = seq.optimizer
    hi = np.asarray(opt.target_link_human_indices)
    names = list(opt.origin_link_names[:1]) + list(opt.task_link_names)
    human = [int(hi[0, 0])] + [int(v) for v in hi[1]]
    return names, human, float(opt.scaling)


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML): a background thread polls every
    2 ms, and the main thread adds one sample right after the launches are enqueued (GPU still busy), so even
    a timed region of a few milliseconds is covered."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.max_mhz, self._h, self._nv = None, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")
        self._t = threading.Thread(target=self._run, daemon=True)

    def sample_now(self):
        if self._h is None:
            return
        nv = self._nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            for bit, nm in self.REASONS.items():
                if r & bit:
                    self.reasons.add(nm)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def _run(self):
        while not self._stop.is_set():
            self.sample_now()
            time.sleep(0.002)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=2)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------- CPU baseline
_WORKER = {}


def _cpu_worker_init():
    from helpers import build_oracle

    _WORKER["o"] = build_oracle(CONFIG_KEY)


def _cpu_worker(args):
    from oracle.solvers import solve_reference

    kp, x0 = args
    o = _WORKER["o"]
    out = []
    for i in range(kp.shape[0]):
        ref = o.ref_from_keypoints(kp[i]).astype(np.float32)
        lastc = np.clip(x0[i], o.joint_limits[:, 0], o.joint_limits[:, 1])
        q, _ = solve_reference(o, ref, np.zeros(0, np.float32), lastc)
        out.append(q)
    return np.array(out)


class CpuReferencePool:
    """The reference-faithful CPU path (oracle mode A: FK / Jacobians in C like the reference's pinocchio, loss in
    numpy, scipy SLSQP at the reference's ftol, value without / gradient with the regulariser) on `cores`
    single-threaded worker processes."""

    def __init__(self, cores, kp, x0):
        import multiprocessing as mp

        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = "1"  # inherited by the spawned workers: one thread each, no oversubscription
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_cpu_worker_init)
        self.pool.map(_cpu_worker, [(kp[:1], x0[:1])] * cores)  # imports + first call outside any timing

    def frames_per_second(self, kp, x0):
        chunks = [(kp[i::self.cores], x0[i::self.cores]) for i in range(self.cores)]
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, chunks)
        return kp.shape[0] / (time.perf_counter() - t0)

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores():
    """Usable host cores: scheduler affinity, capped by the cgroup CPU quota when there is one."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, n)


# --------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from helpers import build_product

    config = {"workload": "VectorOptimizer Allegro right 16-DoF, teleop config (scaling 1.6, huber 0.02, norm_delta 4e-3), "
                          "21-keypoint frames -> qpos, independent frames",
              "frames_per_gpu_per_step": args.frames, "global_frames_per_step": args.frames * world,
              "parallelism": f"frames sharded over {world} GPU(s), no data-path collective",
              "warm_start": "q* + 0.05 N(0,1) clipped", "l2": f"{N_INPUT_SETS} rotating input batches "
              f"({N_INPUT_SETS * args.frames * (BYTES_PER_FRAME - 64) / 1e6:.0f} MB > 126 MB L2)"}

    if args.impl == "reference":
        if rank != 0:
            return
        seq = build_product(CONFIG_KEY)
        cores = host_cores()
        per_step = min(args.frames, max(cores * 64, 64))
        kp, x0 = make_batch(seq.optimizer.robot.kin, workload(seq), per_step * N_INPUT_SETS, 1234)
        pool = CpuReferencePool(cores, kp, x0)
        rates = []
        for s in range(args.warmup + args.steps):
            sl = slice((s % N_INPUT_SETS) * per_step, (s % N_INPUT_SETS + 1) * per_step)
            r = pool.frames_per_second(kp[sl], x0[sl])
            if s >= args.warmup:
                rates.append(r)
        pool.close()
        total = per_step * args.steps
        dt = sum(per_step / r for r in rates)
        value = total / dt
        sample = f"{per_step} frames/step x {args.steps} steps of the same synthetic workload"
        line = {"impl": "reference", "metric": "hand_frames_per_sec", "value": value, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                                 "note": "restated reference path (C FK/Jacobian + numpy loss + scipy SLSQP at the reference's ftol); "
                                         "pinocchio/nlopt are not installable offline"},
                "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    seq = build_product(CONFIG_KEY, device=local_rank)
    opt = seq.optimizer
    if world > 1:
        from dex_retargeting_b200.parallel import broadcast_table

        broadcast_table(opt, src=0)  # the only collective of the path: 8 KB robot table at init
    B = args.frames
    kp_h, x0_h = make_batch(opt.robot.kin, workload(seq), B * N_INPUT_SETS, 1234 + rank)
    kp_sets = [torch.from_numpy(kp_h[i * B:(i + 1) * B]).to(dev) for i in range(N_INPUT_SETS)]
    x0_sets = [torch.from_numpy(x0_h[i * B:(i + 1) * B]).to(dev) for i in range(N_INPUT_SETS)]
    out = torch.empty((B, opt.opt_dof), dtype=torch.float32, device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident arm -------------------------------------------------------------------
    for s in range(args.warmup):
        opt.retarget_batch(keypoints=kp_sets[s % N_INPUT_SETS], last_qpos=x0_sets[s % N_INPUT_SETS], out=out, status_out=status)
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clk:
        evs[0].record()
        for s in range(args.steps):
            opt.retarget_batch(keypoints=kp_sets[s % N_INPUT_SETS], last_qpos=x0_sets[s % N_INPUT_SETS], out=out, status_out=status)
            evs[s + 1].record()
        clk.sample_now()  # launches are enqueued, the GPU is still working through them
        barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    st = status.cpu().numpy()
    iters_mean = float((st & 0xffff).mean())
    flagged = int(((st >> 24) != 0).sum())

    # ---- end-to-end arm: pinned host buffers through the host C-ABI call -------------------------
    kp_pin = [torch.from_numpy(kp_h[i * B:(i + 1) * B]).pin_memory() for i in range(min(2, N_INPUT_SETS))]
    x0_pin = [torch.from_numpy(x0_h[i * B:(i + 1) * B]).pin_memory() for i in range(min(2, N_INPUT_SETS))]
    out_pin = torch.empty((B, opt.opt_dof), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for s in range(2):
        opt.retarget_batch_host(keypoints=kp_pin[s % 2], last_qpos=x0_pin[s % 2], out=out_pin)
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        opt.retarget_batch_host(keypoints=kp_pin[s % 2], last_qpos=x0_pin[s % 2], out=out_pin)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    checksum = float(out_pin.double().sum())

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = json.loads(peaks_path.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        mean_launch_ms = statistics.mean(launch_ms)
        achieved = BYTES_PER_FRAME * B / (mean_launch_ms * 1e-3) / 1e9
        traffic, issue = None, None
        tpath = ROOT / "profiles" / "roofline_traffic.json"
        if tpath.exists():
            tj = json.loads(tpath.read_text())
            traffic = tj.get("dram_bytes_per_launch")
            if tj.get("warp_inst_per_launch"):
                # what actually bounds the solver: warp-instruction issue (4 schedulers/SM, 1 inst/clk each).  Instruction
                # count from the committed ncu capture of this workload, rate from the live launch time and max SM clock.
                sm_mhz = (clk.summary() or {}).get("sm_max_mhz") or 1965
                peak_issue = 4 * torch.cuda.get_device_properties(dev).multi_processor_count * sm_mhz * 1e6
                ach_issue = tj["warp_inst_per_launch"] / (mean_launch_ms * 1e-3)
                issue = {"achieved": ach_issue, "peak": peak_issue, "unit": "warp-inst/s", "frac": ach_issue / peak_issue,
                         "warp_inst_per_frame": tj["warp_inst_per_launch"] / B,
                         "source": "profiles/roofline_traffic.json (ncu smsp__inst_executed.sum) / live launch time"}
        value = B * world * args.steps / (total_ms * 1e-3)
        line = {
            "metric": "hand_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_frame": BYTES_PER_FRAME,
                         "launch_ms_mean": mean_launch_ms, "launch_ms_min": min(launch_ms), "launch_ms_max": max(launch_ms),
                         "note": "latency/FP32-issue bound solver: see `issue` and DESIGN.md for the instruction-level roofline",
                         "issue": issue},
            "e2e": {"value": B * world / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": B * (252 + 64),
                    "d2h_bytes_per_step": B * 64, "ms_per_step": e2e_ms, "steps": e2e_steps, "checksum": checksum,
                    "api": "Optimizer.retarget_batch_host -> dexr_solve_frames_host, pinned host buffers in and out; zero-copy: the "
                           "kernel's TMA producer pulls the input tiles from host memory over PCIe and the results are "
                           "stored straight to host memory, all inside the timed region"},
            "gpu_launches": args.steps, "clocks": clk.summary(),
            "solver": {"mean_iterations": iters_mean, "frames_flagged": flagged, "launch": opt.engine().launch_info()},
        }
        if not args.no_cpu_baseline:
            cores = host_cores()
            n_s = int(min(16384, max(64, cores * 64)))
            pool = CpuReferencePool(cores, kp_h, x0_h)
            r = pool.frames_per_second(kp_h[:n_s], x0_h[:n_s])
            pool.close()
            line["cpu_baseline"] = {"value": r, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n_s} frames of rank 0's first input batch, oracle mode A "
                                              f"(C FK/Jacobian + numpy loss + scipy SLSQP, reference ftol), {cores} processes"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
