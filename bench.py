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
            "visibility": {0: "angular cache (k_heads, k_cam_vis_staged, k_seq, k_finish; ray-march for uncovered envs)",
                           1: "all-fp64 ray-march", 2: "filtered ray-march"}[args.mode],
            "l2": "flushed between timed iterations (256 MiB write)"}


def b_step_bytes(rows, cols, kc, kg):
    """Algorithmic bytes per env-step of the packed step (SURVEY.md 8d / BASELINE.md 4)."""
    g = 4 * rows * ((cols + 31) // 32)
    return 1 + (g + 32 * kc + 40 * kg + 16) + (8 * kc + 12 * kg + 16) + g + 5


def csrc_stamp():
    """sha256 over the kernel sources and the header: ties profiles/traffic.json to the code it was captured from."""
    import hashlib
    h = hashlib.sha256()
    pkg = os.path.join(ROOT, "rl-project-heist-architect-adversarial-reinforcement-learning-framework-cse4019_b200", "csrc")
    for f in sorted(os.listdir(pkg)) + ["../../include/heist_b200.h"]:
        h.update(open(os.path.join(pkg, f), "rb").read())
    return h.hexdigest()[:16]


def measured(key, field):
    """Per-step figures REPLAYED from the committed ncu capture of this command (profiles/traffic.json, written by
    profiles/make_traffic.py), or None when there is none or it was taken from other kernel sources (its `csrc`
    stamp differs from csrc_stamp()): dram_bytes_per_launch = dram__bytes_read.sum + dram__bytes_write.sum,
    warp_inst_per_launch = smsp__inst_executed.sum, summed over the kernels of one step."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
        return d[field] if d.get("csrc") == csrc_stamp() else None
    except Exception:
        return None


def traffic_status(key):
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
    except Exception:
        return "no capture committed"
    return "replayed from profiles/traffic.json (ncu capture of this command at these kernel sources)" if d.get("csrc") == csrc_stamp() \
        else "null: profiles/traffic.json was captured from other kernel sources (stale)"


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
PY_REFERENCE = {"steps_per_s_per_core": 19.8, "what": "unmodified HeistEnvironment.step, 20x20, 2 cameras + 1 guard + 4 walls",
                "where": "survey container (1 host core, CPython 3.12, numpy 2.3); the Python reference cannot travel to the GPU box",
                "source": "BASELINE.md section 3"}


class Timer:
    """W warm-up + exactly K timed iterations of `body`, L2 flushed (untimed) before each, CUDA events on the
    caller's stream, barrier + synchronize on both sides, MAX over ranks."""

    def __init__(self, torch, dist, dev, world, flush):
        self.torch, self.dist, self.dev, self.world, self.flush = torch, dist, dev, world, flush

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, body, warmup, steps):
        torch = self.torch
        for i in range(warmup):
            self.flush.fill_(i & 0xFF)
            body(i)
        self.barrier()
        evs = []
        for i in range(steps):
            self.flush.fill_(i & 0xFF)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            body(warmup + i)
            e.record()
            evs.append((s, e))
        self.barrier()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def rollout_leg(ctx, wl, steps, warmup, want_e2e, want_layout=False):
    """One rollout workload (a BASELINE config) on this rank: build the env, sample valid layouts, time step_many.
    Returns a dict of raw timings; the env is closed before returning."""
    import numpy as np
    torch, hb, synthetic, timer = ctx["torch"], ctx["hb"], ctx["synthetic"], ctx["timer"]
    dev, rank = ctx["dev"], ctx["rank"]
    cfg = hb.EnvironmentConfig(grid_rows=wl["rows"], grid_cols=wl["cols"], max_steps=200, architect_budget=wl["budget"])
    env = hb.BatchedHeistEnv(cfg, wl["envs"], device=dev)
    env.set_mode(wl.get("mode", 0))
    seed = synthetic.BASE_SEED + rank
    am_host, cp_host = synthetic.make_valid_workload(env, seed, wl["budget"], exact_counts=wl.get("exact_counts"))
    env.reset()
    envs_cached, cache_bytes = env.cache_stats()
    kc = float(env.env_static[:, 0].float().mean().item())
    kg = float(env.env_static[:, 1].float().mean().item())
    rng = np.random.default_rng(seed + 7919)
    T, N = wl["ticks"], wl["envs"]
    n_iter = warmup + steps
    # one action tensor per iteration, resident in HBM before the timed region (value) and in pinned host memory (e2e)
    acts_host = [torch.from_numpy(synthetic.sample_actions(rng, T, N)) for _ in range(n_iter)]
    acts_dev = [a.to(dev) for a in acts_host]
    # every env step writes its packed observable state to HBM: reward, done, status and the visibility bitmap
    out = {"reward": torch.empty((T, N), dtype=torch.float32, device=dev),
           "done": torch.empty((T, N), dtype=torch.uint8, device=dev),
           "status": torch.empty((T, N), dtype=torch.uint8, device=dev),
           "vis_bits": torch.empty((T, N, env.R, env.W), dtype=torch.int32, device=dev)}
    launches0 = env.launch_count()
    ms_value = timer.run(lambda i: env.step_many(acts_dev[i], autoreset=True, out=out), warmup, steps)
    res = {"ms_value": ms_value, "launches_per_step": (env.launch_count() - launches0) // n_iter, "kc": kc, "kg": kg,
           "envs_cached": envs_cached, "cache_bytes": cache_bytes, "steps_per_iter": T * N, "cache_warning": env.cache_warning}
    if want_e2e:   # end to end through the public API with HOST buffers: H2D actions, rollout, D2H reward + done
        acts_pin = [a.pin_memory() for a in acts_host]
        host_out = {"reward": torch.empty((T, N), dtype=torch.float32).pin_memory(),
                    "done": torch.empty((T, N), dtype=torch.uint8).pin_memory()}
        res["ms_e2e"] = timer.run(lambda i: env.step_many_host(acts_pin[i], host_out, autoreset=True, vis_out=out["vis_bits"]),
                                  warmup, steps)
    if want_layout and rank == 0:
        res["extra"] = secondary_kernels(ctx, env, wl, out, am_host, cp_host)
    env.close()
    del env, out, acts_dev
    torch.cuda.empty_cache()
    return res


def secondary_kernels(ctx, env, wl, out, am_host, cp_host):
    """The other kernels of the path, timed alone (reported, not the headline): dense observation, GAE, the fused
    single tick, and the per-layout work (decode + placement + BFS + visibility tables)."""
    torch, hb, flush = ctx["torch"], ctx["hb"], ctx["timer"].flush
    dev = ctx["dev"]
    T, N, R, C = wl["ticks"], wl["envs"], wl["rows"], wl["cols"]
    peak = ctx["peak"]

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

    state = torch.empty((N, 3, R, C), dtype=torch.float32, device=dev)
    ms_obs = time_kernel(lambda: env.observe(out=state))
    g = 4 * R * ((C + 31) // 32)
    obs_bytes = N * (12 * R * C + g + R * C + 4)
    val = torch.randn((T, N), device=dev)
    ms_gae = time_kernel(lambda: hb.compute_gae(out["reward"], val, out["done"]))
    a1 = torch.zeros(N, dtype=torch.int8, device=dev)
    ms_tick = time_kernel(lambda: env.step_observe(a1, autoreset=True, state_out=state))

    def graphed_us(fn, n=20, reps=10):
        """Device time per call when n calls replay from one CUDA graph (how a policy loop drives single ticks:
        no host launch gaps, tables warm in L2)."""
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                fn()
        g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        return 1e3 * s.elapsed_time(e) / (reps * n)

    try:
        env.set_mode(env.MODE_TABLES)   # every env of the workload is table-driven: the tick is exactly one kernel
        us_tick_graph = graphed_us(lambda: env.step_observe(a1, autoreset=True, state_out=state))
        env.set_mode(wl.get("mode", 0))
        us_gae_graph = graphed_us(lambda: hb.compute_gae(out["reward"], val, out["done"]))
    except Exception as ex:   # reported, never fatal for the headline
        us_tick_graph = us_gae_graph = None
        sys.stderr.write(f"[bench] graph timing skipped: {ex}\n")
    am_dev, cp_dev = torch.as_tensor(am_host).to(dev), torch.as_tensor(cp_host).to(dev)
    ms_layout = time_kernel(lambda: env.set_layout_from_asset_map(am_dev, cp_dev, wl["budget"]), reps=5)
    env.reset()
    return {"layout": {"ms": ms_layout, "us_per_env": 1e3 * ms_layout / N,
                       "what": "heist_decode_validate: k_decode + k_set_layout (BFS) + k_build_cache + k_build_order, once per layout"},
            "observe": {"ms": ms_obs, "achieved_gbs": obs_bytes / ms_obs / 1e6, "frac": obs_bytes / ms_obs / 1e6 / peak, "bytes": obs_bytes},
            "gae": {"ms": ms_gae, "achieved_gbs": 17 * T * N / ms_gae / 1e6, "frac": 17 * T * N / ms_gae / 1e6 / peak, "bytes": 17 * T * N},
            "step_observe_tick": {"us": 1e3 * ms_tick, "what": "heist_step_observe: one tick + auto-reset + dense state for all envs (one fused "
                                  "kernel); us = eager launch after an L2 flush, us_graph = per tick when 20 ticks replay from a CUDA graph (HEIST_MODE_TABLES)",
                                  "us_graph": us_tick_graph, "achieved_gbs": obs_bytes / ms_tick / 1e6, "frac": obs_bytes / ms_tick / 1e6 / peak,
                                  "frac_graph": (obs_bytes / (us_tick_graph * 1e-3) / 1e6 / peak) if us_tick_graph else None},
            "gae_graph_us": us_gae_graph}


def ppo_leg(ctx, envs, ticks, iters, warmup):
    """BASELINE config 5: the full adversarial loop (heist_b200.loop.AdversarialLoop) with networks of the reference's
    shapes, one process per GPU, NCCL gradient all-reduce.  Device-timed; MAX over ranks."""
    torch, hb, dist = ctx["torch"], ctx["hb"], ctx["dist"]
    dev, rank, world = ctx["dev"], ctx["rank"], ctx["world"]
    from heist_b200 import nets
    from heist_b200.loop import AdversarialLoop
    torch.manual_seed(4242)   # identical initial weights on every rank
    solver, architect = nets.SolverNet().to(dev), nets.ArchitectNet().to(dev)
    torch.manual_seed(4242 + rank)
    env = hb.BatchedHeistEnv(hb.EnvironmentConfig(), envs, device=dev)
    loop = AdversarialLoop(env, solver, architect, ticks=ticks, budget=15, group=None)
    phases, tot = {}, 0.0
    for it in range(warmup + iters):
        _, ms = loop.iteration(temperature=1.0)
        if it >= warmup:
            for k, v in ms.items():
                phases[k] = phases.get(k, 0.0) + v / iters
    t = torch.tensor([phases[k] for k in sorted(phases)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    phases = dict(zip(sorted(phases), [float(x) for x in t.tolist()]))
    ar = loop.time_allreduce()
    n_updates = loop.epochs * (-(-(ticks * envs) // loop.minibatch))
    res = {"workload": f"config5: full adversarial PPO loop, 20x20, {envs} layouts/GPU x {ticks} ticks per iteration, policy in the loop, "
                       f"SolverNet 550150 / ArchitectNet 407464 parameters (reference shapes, random init), {loop.epochs} epochs x "
                       f"minibatch {loop.minibatch}, fp32 (TF32 convolutions)",
           "n_gpus": world, "value": ticks * envs * world / (phases["total"] * 1e-3), "unit": "env-steps/s (policy in the loop, whole job)",
           "ms_per_ppo_iteration": phases["total"], "ms_phases": {k: round(v, 3) for k, v in phases.items() if k != "total"},
           "us_per_tick_with_policy": 1e3 * phases["rollout"] / ticks, "tick_graph_captured": loop.graphed,
           "optimizer_steps_per_iteration": n_updates,
           "allreduce": {"solver_us": ar["solver_us"], "architect_us": ar["architect_us"], "solver_bytes": ar["solver_bytes"],
                         "architect_bytes": ar["architect_bytes"], "per_iteration_ms": (n_updates * ar["solver_us"] + ar["architect_us"]) * 1e-3,
                         "how": "persistent flat bucket (GradBucket), NCCL ReduceOp.AVG, issued async and overlapped with the next "
                                "minibatch's heist_expand_states; timed alone, device events"}}
    if not loop.graphed:
        res["tick_graph_error"] = getattr(loop, "graph_error", "")
    env.close()
    del loop, env
    torch.cuda.empty_cache()
    return res


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import heist_b200
    from heist_b200 import synthetic

    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak, peak_src = hbm_peak()
    ctx = {"torch": torch, "dist": dist, "hb": heist_b200, "synthetic": synthetic, "dev": dev, "rank": rank, "world": world,
           "timer": Timer(torch, dist, dev, world, flush), "peak": peak}
    wl = {"rows": args.rows, "cols": args.cols, "envs": args.envs, "budget": args.budget, "exact_counts": args.exact_counts,
          "ticks": args.ticks, "mode": args.mode}

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # (1) kernel-resident throughput (inputs already in HBM) and (2) end to end with host buffers
    r = rollout_leg(ctx, wl, args.steps, args.warmup, want_e2e=True, want_layout=True)
    clocks = sampler.stop() if sampler else None
    T, N = args.ticks, args.envs
    steps_per_iter = T * N  # auto-reset: every (tick, env) is a live env step
    total_steps = steps_per_iter * args.steps * world
    value = total_steps / (r["ms_value"] * 1e-3)
    e2e_value = total_steps / (r["ms_e2e"] * 1e-3)

    # (3) the other BASELINE configs under the same clock: 3 = stress (rank 0), 4 = 262144 envs 64x64 SHARDED over the
    #     ranks (strong scaling), 5 = the full adversarial PPO loop with the NCCL gradient all-reduce
    legs = {}
    if args.legs and args.config == 2:
        if rank == 0:
            w3 = {"rows": 32, "cols": 32, "envs": 65536, "budget": 22, "exact_counts": (2, 4, 2), "ticks": 200, "mode": args.mode}
            t3 = Timer(torch, dist, dev, 1, flush)
            r3 = rollout_leg({**ctx, "timer": t3, "world": 1}, w3, 3, 3, want_e2e=False, want_layout=True)
            b3 = b_step_bytes(32, 32, r3["kc"], r3["kg"])
            legs["config3"] = {"workload": "config3: 32x32 grid, 65536 envs on one GPU, budget 22 (4 cameras + 2 guards + 2 walls), T=200, auto-reset",
                               "value": r3["steps_per_iter"] * 3 / (r3["ms_value"] * 1e-3), "unit": UNIT, "n_gpus": 1,
                               "ms_per_step": r3["ms_value"] / 3, "steps": 3, "warmup": 3,
                               "hbm_frac": b3 * r3["steps_per_iter"] * 3 / (r3["ms_value"] * 1e-3) / 1e9 / peak,
                               "bytes_per_env_step": b3, "cache_bytes": r3["cache_bytes"], "envs_cached": r3["envs_cached"],
                               "other_kernels": r3.get("extra", {})}
        ctx["timer"].barrier()
        n4 = 262144 // world
        w4 = {"rows": 64, "cols": 64, "envs": n4, "budget": 22, "exact_counts": (2, 4, 2), "ticks": 200, "mode": args.mode}
        r4 = rollout_leg(ctx, w4, 3, 3, want_e2e=False)
        b4 = b_step_bytes(64, 64, r4["kc"], r4["kg"])
        legs["config4"] = {"workload": f"config4: 64x64 grid, 262144 envs sharded over {world} GPU(s) ({n4} per GPU), budget 22, T=200, auto-reset",
                           "value": r4["steps_per_iter"] * 3 * world / (r4["ms_value"] * 1e-3), "unit": UNIT, "n_gpus": world,
                           "scaling": "strong", "ms_per_step": r4["ms_value"] / 3, "steps": 3, "warmup": 3, "envs_per_gpu": n4,
                           "hbm_frac_per_gpu": b4 * r4["steps_per_iter"] * 3 / (r4["ms_value"] * 1e-3) / 1e9 / peak,
                           "bytes_per_env_step": b4, "cache_bytes_per_gpu": r4["cache_bytes"], "envs_cached_per_gpu": r4["envs_cached"]}
        legs["config5"] = ppo_leg(ctx, 4096, 64, 2, 1)

    if rank != 0:
        return
    extra = r.get("extra", {})
    ms_kernel = r["ms_value"] / args.steps
    cache_bytes, envs_cached, kc, kg = r["cache_bytes"], r["envs_cached"], r["kc"], r["kg"]
    launches_per_step = r["launches_per_step"]
    tkey = "step_cached" if (args.mode == 0 and cache_bytes) else "k_step_many"
    bstep = b_step_bytes(args.rows, args.cols, kc, kg)
    achieved = bstep * steps_per_iter / (ms_kernel * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_kernel, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": T * N, "d2h_bytes_per_step": 5 * T * N,
                    "ms_per_step": r["ms_e2e"] / args.steps,
                    "note": "pinned host actions in (1 B / env-step), pinned host reward + done out (5 B / env-step); status, the "
                            "visibility bitmaps (80 B / env-step at 20x20) and observations stay on the device for an on-device "
                            "policy -- a host-side consumer of the bitmaps would be PCIe-bound near 6e8 env-steps/s"},
            "gpu_launches": launches_per_step * args.steps,  # counted by the library (heist_launch_count)
            "roofline": {"bound": "hbm", "kernel": tkey, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured(tkey, "dram_bytes_per_launch"), "traffic_status": traffic_status(tkey),
                         "peak_source": peak_src, "bytes_per_env_step": bstep, "mean_cams": kc, "mean_guards": kg,
                         "launches_per_step": launches_per_step, "kernel_share": measured(tkey, "share"),
                         "note": "achieved = SURVEY 8d algorithmic bytes of the packed step x env-steps / duration of the whole "
                                 "step (all of its kernels, pipelined over several streams, timed on the caller's stream); the "
                                 "step is bound by table lookups and warp-instruction issue, not by HBM (DESIGN.md 4): frac is "
                                 "honest and small, traffic is the step's DRAM bytes from the committed ncu capture"},
            "cache": {"envs_cached": envs_cached, "envs": N, "bytes": cache_bytes, "warning": r["cache_warning"]},
            "clocks": clocks, "other_kernels": extra}
    if "layout" in extra:   # env-steps/s with the per-layout work charged: one set_layout per solver_attempts x max_steps ticks
        per_layout = 20 * 200   # training.py:515 (solver_attempts = 20) x max_steps
        ms_amort = ms_kernel * (per_layout / T) + extra["layout"]["ms"]
        line["amortised_value"] = {"value": per_layout * N * world / (ms_amort * 1e-3), "unit": UNIT,
                                   "what": "value with heist_decode_validate (decode + BFS + visibility tables) charged once per "
                                           "20 attempts x 200 ticks, the reference trainer's ratio (training.py:515)",
                                   "layout_ms": extra["layout"]["ms"]}
    inst = measured(tkey, "warp_inst_per_launch")
    if inst and args.config == 2 and args.envs == 4096 and args.ticks == 200 and clocks and clocks.get("sm_mhz"):
        # informational: the bound that actually applies.  Warp-instructions per launch from the committed ncu
        # capture, issue peak = 148 SMs x 4 schedulers x 1 instruction/clock at the SM clock sampled during the run.
        peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6
        ach = inst / (ms_kernel * 1e-3)
        line["roofline_issue"] = {"bound": "warp-instruction issue", "achieved": ach, "peak": peak_issue,
                                  "unit": "warp-inst/s", "frac": ach / peak_issue,
                                  "warp_inst_per_env_step": inst / steps_per_iter, "source": traffic_status(tkey)}
    if legs:
        line["other_configs"] = legs
    if not args.no_cpu_baseline and world == 1:
        rate, threads, sample, _, _ = cpu_rollout_rate(args, 256, args.cpu_seconds, synthetic.BASE_SEED + rank)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                                "python_reference": PY_REFERENCE}
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
    ap.add_argument("--no-legs", dest="legs", action="store_false",
                    help="skip the other BASELINE configs (3, 4 sharded, 5 full PPO loop) that the default run adds as other_configs")
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
