#!/usr/bin/env python
"""bench.py -- Solver env-steps/sec of the batched Heist environment hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: a T-tick rollout (heist_step_many, auto-reset,
pre-generated uniform actions) of the per-GPU env batch.  Workload at every N: BASELINE config 2
(20x20 grid, 4096 envs per GPU, random valid layouts decoded from sampled asset maps at budget 15,
max_steps 200) -> weak scaling, no data-path collective.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "solver_env_steps_per_sec"
UNIT = "env-steps/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def workload_config(args, world):
    return {"workload": f"config{args.config}: {args.rows}x{args.cols} grid, {args.envs} envs/GPU, random valid layouts "
                        f"(budget {args.budget}{', walls/cameras/guards = %d/%d/%d' % args.exact_counts if args.exact_counts else ''}), "
                        f"Solver-only rollout T={args.ticks}, auto-reset, uniform actions",
            "grid": [args.rows, args.cols], "envs_per_gpu": args.envs, "ticks_per_step": args.ticks,
            "budget": args.budget, "max_steps": 200, "parallelism": f"env-shard x{world} (no data-path collective)",
            "visibility": {0: "angular cache (k_heads, k_cam_vis, k_seq, k_finish; ray-march for uncovered envs)",
                           1: "all-fp64 ray-march", 2: "filtered ray-march"}[args.mode],
            "l2": "flushed between timed iterations (256 MiB write)"}


def b_step_bytes(rows, cols, kc, kg):
    """Algorithmic bytes per env-step of the packed step (SURVEY.md 8d / BASELINE.md 4)."""
    g = 4 * rows * ((cols + 31) // 32)
    return 1 + (g + 32 * kc + 40 * kg + 16) + (8 * kc + 12 * kg + 16) + g + 5


def measured(key, field):
    """Per-step figures from the committed ncu capture of this command (profiles/traffic.json), or None:
    dram_bytes_per_launch = dram__bytes_read.sum + dram__bytes_write.sum, warp_inst_per_launch = smsp__inst_executed.sum,
    summed over the kernels of one step."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key][field]
    except Exception:
        return None


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.rows = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU side: oracle port (test infrastructure) as the reported baseline / reference arm
# ------------------------------------------------------------------------------------------------
def cpu_rollout_rate(args, n_envs, budget_s, seed, n_threads=0, min_rounds=1):
    """Time the oracle port on a bounded sample of the same workload. Returns (steps/s, threads, sample, rounds)."""
    import numpy as np
    from heist_b200 import synthetic
    from oracle import heist_oracle as ho
    rng = np.random.default_rng(seed)
    envs = []
    while len(envs) < n_envs:  # random VALID layouts: resample invalid ones (validity from the oracle's BFS)
        am = (synthetic.sample_asset_maps_exact(rng, 1, args.rows, args.cols, *args.exact_counts)[0]
              if args.exact_counts else synthetic.sample_asset_maps(rng, 1, args.rows, args.cols)[0])
        cp = synthetic.sample_cam_params(rng, 1)[0]
        walls, cams, guards, _ = ho.decode_layout(am, args.budget, *cp)
        e = ho.OracleEnv(args.rows, args.cols, max_steps=200, budget=args.budget)
        if e.set_layout(walls, cams, guards):
            envs.append(e)
    ho.reset_all(envs, n_threads)
    threads = n_threads if n_threads > 0 else ho.num_threads()
    steps, elapsed, rounds = 0, 0.0, 0
    while rounds < min_rounds or elapsed < budget_s:
        acts = synthetic.sample_actions(rng, args.ticks, n_envs)
        t0 = time.perf_counter()
        out = ho.rollout(envs, acts, autoreset=True, want_vis=False, n_threads=n_threads)
        elapsed += time.perf_counter() - t0
        steps += int(out["live"])
        rounds += 1
    sample = f"{n_envs} envs x {args.ticks} ticks x {rounds} rounds of the same workload (seed {seed})"
    return steps / elapsed, threads, sample, rounds, elapsed


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure Python and
    is not present on the GPU box, so this is the oracle port (C restatement pinned to the reference through
    tests/golden), on all host threads.  Rank 0 only."""
    if rank != 0:
        return
    n_envs = min(args.envs, 512)
    # warm-up rounds then K timed rounds, each a bounded sample
    cpu_rollout_rate(args, n_envs, 0.0, 1, min_rounds=max(1, min(args.warmup, 2)))
    rate, threads, sample, rounds, elapsed = cpu_rollout_rate(args, n_envs, 0.0, 2, min_rounds=args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / rounds,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import heist_b200
    from heist_b200 import synthetic

    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    cfg = heist_b200.EnvironmentConfig(grid_rows=args.rows, grid_cols=args.cols, max_steps=200,
                                       architect_budget=args.budget)
    env = heist_b200.BatchedHeistEnv(cfg, args.envs, device=dev)
    env.set_mode(args.mode)
    seed = synthetic.BASE_SEED + rank
    am_host, cp_host = synthetic.make_valid_workload(env, seed, args.budget, exact_counts=args.exact_counts)
    env.reset()
    envs_cached, cache_bytes = env.cache_stats()
    kc = float(env.env_static[:, 0].float().mean().item())
    kg = float(env.env_static[:, 1].float().mean().item())
    rng = np.random.default_rng(seed + 7919)
    T, N = args.ticks, args.envs
    n_iter = args.warmup + args.steps
    # one action tensor per iteration, resident in HBM before the timed region (value) and in pinned host
    # memory (e2e)
    acts_host = [torch.from_numpy(synthetic.sample_actions(rng, T, N)).pin_memory() for _ in range(n_iter)]
    acts_dev = [a.to(dev) for a in acts_host]
    # every env step writes its packed observable state to HBM: reward, done, status and the visibility bitmap
    out = {"reward": torch.empty((T, N), dtype=torch.float32, device=dev),
           "done": torch.empty((T, N), dtype=torch.uint8, device=dev),
           "status": torch.empty((T, N), dtype=torch.uint8, device=dev),
           "vis_bits": torch.empty((T, N, env.R, env.W), dtype=torch.int32, device=dev)}
    rew_host = torch.empty((T, N), dtype=torch.float32).pin_memory()
    done_host = torch.empty((T, N), dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(body):
        """W warm-up + exactly K timed iterations; L2 flushed (untimed) before each; device-timed."""
        for i in range(args.warmup):
            flush.fill_(i & 0xFF)
            body(i)
        barrier()
        evs = []
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            body(args.warmup + i)
            e.record()
            evs.append((s, e))
        barrier()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()

    # (1) kernel-resident throughput: inputs already in HBM
    launches0 = env.launch_count()
    ms_value = timed_loop(lambda i: env.step_many(acts_dev[i], autoreset=True, out=out))
    launches_per_step = (env.launch_count() - launches0) // (args.warmup + args.steps)
    steps_per_iter = T * N  # auto-reset: every (tick, env) is a live env step
    total_steps = steps_per_iter * args.steps * world
    value = total_steps / (ms_value * 1e-3)

    # (2) end to end through the public API with HOST buffers: H2D actions, rollout, D2H reward + done
    host_out = {"reward": rew_host, "done": done_host}

    def e2e_body(i):   # pinned host actions in, pinned host reward + done out (the copies ride along the launch)
        env.step_many_host(acts_host[i], host_out, autoreset=True, vis_out=out["vis_bits"])

    ms_e2e = timed_loop(e2e_body)
    e2e_value = total_steps / (ms_e2e * 1e-3)
    clocks = sampler.stop() if sampler else None

    # (3) secondary HBM-streaming kernels of the path, timed alone (reported, not the headline)
    def time_kernel(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for i in range(reps):
            flush.fill_(i & 0xFF)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        return tot / reps

    peak, peak_src = hbm_peak()
    extra = {}
    if rank == 0:
        state = torch.empty((N, 3, args.rows, args.cols), dtype=torch.float32, device=dev)
        ms_obs = time_kernel(lambda: env.observe(out=state))
        g = 4 * args.rows * ((args.cols + 31) // 32)
        obs_bytes = N * (12 * args.rows * args.cols + g + args.rows * args.cols + 4)
        val = torch.randn((T, N), device=dev)
        ms_gae = time_kernel(lambda: heist_b200.compute_gae(out["reward"], val, out["done"]))
        # per-layout work, outside the timed step: decode + placement + BFS + visibility tables + slot order
        am_dev, cp_dev = torch.as_tensor(am_host).to(dev), torch.as_tensor(cp_host).to(dev)
        ms_layout = time_kernel(lambda: env.set_layout_from_asset_map(am_dev, cp_dev, args.budget), reps=5)
        env.reset()
        extra = {"layout": {"ms": ms_layout, "us_per_env": 1e3 * ms_layout / N,
                            "what": "heist_decode_validate: k_decode + k_set_layout (BFS) + k_build_cache + k_build_order, once per layout"},
                 "observe": {"ms": ms_obs, "achieved_gbs": obs_bytes / ms_obs / 1e6, "frac": obs_bytes / ms_obs / 1e6 / peak,
                             "bytes": obs_bytes},
                 "gae": {"ms": ms_gae, "achieved_gbs": 17 * T * N / ms_gae / 1e6, "frac": 17 * T * N / ms_gae / 1e6 / peak,
                         "bytes": 17 * T * N}}

    if rank != 0:
        return
    ms_kernel = ms_value / args.steps
    tkey = "step_cached" if (args.mode == 0 and cache_bytes) else "k_step_many"
    bstep = b_step_bytes(args.rows, args.cols, kc, kg)
    achieved = bstep * steps_per_iter / (ms_kernel * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_kernel, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": T * N, "d2h_bytes_per_step": 5 * T * N,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,  # counted by the library (heist_launch_count)
            "roofline": {"bound": "hbm", "kernel": tkey, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured(tkey, "dram_bytes_per_launch"), "peak_source": peak_src,
                         "bytes_per_env_step": bstep, "mean_cams": kc, "mean_guards": kg,
                         "launches_per_step": launches_per_step, "kernel_share": measured(tkey, "share"),
                         "note": "achieved = SURVEY 8d algorithmic bytes of the packed step x env-steps / duration of the whole "
                                 "step (all of its kernels, pipelined over three streams, timed on the caller's stream); the "
                                 "step is bound by L2-resident table lookups and warp-instruction issue, not by HBM "
                                 "(DESIGN.md 4): frac is honest and small, traffic is the step's measured DRAM bytes"},
            "cache": {"envs_cached": envs_cached, "envs": N, "bytes": cache_bytes},
            "clocks": clocks, "other_kernels": extra}
    inst = measured(tkey, "warp_inst_per_launch")
    if inst and args.config == 2 and args.envs == 4096 and args.ticks == 200 and clocks and clocks.get("sm_mhz"):
        # informational: the bound that actually applies.  Warp-instructions per launch from the committed ncu
        # capture, issue peak = 148 SMs x 4 schedulers x 1 instruction/clock at the SM clock sampled during the run.
        peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6
        ach = inst / (ms_kernel * 1e-3)
        line["roofline_issue"] = {"bound": "warp-instruction issue", "achieved": ach, "peak": peak_issue,
                                  "unit": "warp-inst/s", "frac": ach / peak_issue,
                                  "warp_inst_per_env_step": inst / steps_per_iter}
    if not args.no_cpu_baseline and world == 1:
        rate, threads, sample, _, _ = cpu_rollout_rate(args, 256, args.cpu_seconds, seed)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 100 for config 2, 5 / 3 for configs 3 / 4)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--rows", type=int, default=20)
    ap.add_argument("--cols", type=int, default=20)
    ap.add_argument("--ticks", type=int, default=200)
    ap.add_argument("--budget", type=int, default=15)
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4],
                    help="BASELINE.json configs[] index: 2 = 20x20/4096 envs (default, the headline), 3 = 32x32/65536 "
                         "envs, 4 cameras + 2 guards, 4 = 64x64/262144 envs split over the GPUs")
    ap.add_argument("--mode", type=int, default=0, choices=[0, 1, 2],
                    help="heist_set_mode: 0 = angular visibility cache (default), 1 = all-fp64 ray-march, 2 = filtered ray-march")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {2: 100, 3: 5, 4: 3}[args.config] if args.impl == "ours" else 5
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.exact_counts = None
    if args.config == 3:
        args.rows = args.cols = 32
        args.envs, args.budget, args.exact_counts = 65536, 22, (2, 4, 2)
    elif args.config == 4:
        args.rows = args.cols = 64
        args.envs, args.budget, args.exact_counts = 262144 // max(1, int(os.environ.get("WORLD_SIZE", "1"))), 22, (2, 4, 2)

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl == "ours":
        # convenience: the driver launches torchrun itself; a bare `python bench.py --gpus N` re-launches under it
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
                                   "29533", os.path.abspath(__file__)] + sys.argv[1:])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        # NCCL writes its version banner (NCCL_DEBUG=VERSION/INFO on some boxes) to stdout: keep stdout for the one JSON
        # line by pointing fd 1 at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local_rank}"))
            dist.all_reduce(torch.zeros(1, device=f"cuda:{local_rank}"))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
