#!/usr/bin/env python
"""bench.py -- Azul env steps/s of the batched random-agent rollout (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One bench "step" = one pass of the hot path over the batch: ONE launch of the fused rollout kernel
(azb_rollout_random: legal mask + random agent + Azul.step + scoring + refill + auto-reset) that
advances every one of the G games by --k-steps env steps.  N = 1 runs BASELINE.json configs[1]
(65,536 parallel 2-player random-agent games); N > 1 (torchrun, one rank per GPU) shards the global
game-id range over the ranks with no data-path collective (weak scaling, G games per GPU) and uses
NCCL only for the max-over-ranks timing and the rollout-counter reduction.

The JSON line carries
  value      env steps/s, state resident in HBM, timed with CUDA events on the launching stream
  e2e        the same metric through the public host API with HOST buffers: per step the packed state
             is copied from pinned host memory to the device, rolled out, and state + legal mask +
             counters are copied back
  roofline   algorithmic bytes (2*S(P)+25 per env step) / kernel time vs measured HBM peak
  cpu_baseline  the C oracle port timed on this host's cores on a bounded sample (N = 1, rank 0)
--impl reference times the reference arm for this tier: the oracle port (the reference is pure Python
and cannot travel to the GPU box) on all host threads, same config / metric / unit.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "azul_env_steps_per_sec"
UNIT = "env_steps/s"
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="random", choices=["random", "policy", "train", "step"],
                    help="random: fused random-agent rollout (configs[1], the headline); policy: self-play with the "
                         "fused ActorCritic policy kernel, one env step per launch (configs[3])")
    ap.add_argument("--games", type=int, default=0, help="games per GPU (default 65536 random / 131072 policy)")
    ap.add_argument("--players", type=int, default=2)
    ap.add_argument("--pool", default="lid", choices=["lid", "random"],
                    help="tile pool; 'lid' + random first player = GameRunner's default rules (game_runner.py:23)")
    ap.add_argument("--k-steps", type=int, default=4096,
                    help="env steps per game per launch (SURVEY §8d config 2: K = 4,096 with auto-reset)")
    ap.add_argument("--block", type=int, default=0, help="threads per block (0 = library default)")
    ap.add_argument("--defer", type=int, default=0, help="rollout end-of-round batching threshold (0 = library default)")
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0x5EED)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--step-actions", default="random", choices=["random", "lowest"],
                    help="--mode step: the caller-supplied action per game (a random or the lowest legal one)")
    ap.add_argument("--presteps", type=int, default=203,
                    help="--mode step: random-agent env steps before the timed step (0 = fresh games: no round can end)")
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {
        "workload": "%d parallel %d-player random-agent Azul games per GPU, env step + legal mask, auto-reset "
                    "(BASELINE.json configs[1])" % (args.games, args.players),
        "games_per_gpu": args.games, "players": args.players,
        "rules": {"tile_pool": "Lid" if args.pool == "lid" else "Random", "first_player": "Random"},
        "env_steps_per_game_per_launch": args.k_steps,
        "rng": "Philox4x32-10, seed 0x%X, keyed by global game id" % args.seed,
        "l2": "flushed between timed launches (256 MiB device fill outside the CUDA-event brackets)",
        "parallelism": "games sharded by global id over %d GPU(s); no data-path collective" % n_gpus,
    }


# ------------------------------------------------------------------------------------------
# CPU legs (oracle port): cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------
def cpu_rollout_rate(args, seconds, threads):
    """Time the C oracle's rollout (same Philox schedule, same rules) on `threads` host threads."""
    from oracle import oracle as O
    pool = 1 if args.pool == "lid" else 0
    probe_n, probe_k = 256 * threads, 64
    recs = O.fresh_records(probe_n, args.players, pool, 0, args.seed, 0)
    t0 = time.perf_counter()
    O.rollout_random(recs, args.players, pool, 0, args.seed, 0, probe_k, threads=threads)
    rate = probe_n * probe_k / max(time.perf_counter() - t0, 1e-6)
    k = args.k_steps
    n = int(max(threads, min(args.games, rate * seconds / k)))
    recs = O.fresh_records(n, args.players, pool, 0, args.seed, 0)
    t0 = time.perf_counter()
    cnt = O.rollout_random(recs, args.players, pool, 0, args.seed, 0, k, threads=threads)
    dt = time.perf_counter() - t0
    assert cnt[0] == n * k
    return n * k / dt, n, k, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                   # rank 0 alone runs the CPU arm
    threads = os.cpu_count() or 1
    pool = 1 if args.pool == "lid" else 0
    from oracle import oracle as O
    # bounded sample per step so the whole run ends within minutes: ~1.5 s of CPU work per step
    rate, _, _, _ = cpu_rollout_rate(args, 1.0, threads)
    n = int(max(threads, min(args.games, rate * 1.5 / args.k_steps)))
    recs = O.fresh_records(n, args.players, pool, 0, args.seed, 0)
    for _ in range(args.warmup):
        O.rollout_random(recs, args.players, pool, 0, args.seed, 0, args.k_steps, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.rollout_random(recs, args.players, pool, 0, args.seed, 0, args.k_steps, threads=threads)
    dt = time.perf_counter() - t0
    value = n * args.k_steps * args.steps / dt
    sample = "%d of %d games x %d env steps per step, C oracle port (oracle/azul_oracle.c), %d threads" % (
        n, args.games, args.k_steps, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is pure Python and cannot travel to the GPU box; this arm is the C restatement "
                "pinned bit-exactly to it (tests/test_oracle_golden.py). Python reference measured in the build "
                "container: ~3.4e3 env steps/s/core (BASELINE.md §2).",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled every 10 ms through NVML (nvidia_ml_py) while the timed region
    runs; falls back to an `nvidia-smi -lms` subprocess (the profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, torch_device_index):
        import threading
        self.idx, self.proc, self.thread, self.stop_flag = torch_device_index, None, None, threading.Event()
        self.sm, self.bits, self.max_mhz, self.power = [], 0, None, []
        self.nvml = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _loop(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self.stop_flag.wait(0.01)

    def start(self):
        if self.nvml is not None:
            import threading
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join()
            reasons = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None,
                    "reasons": reasons, "source": "nvml, 10 ms period, timed region only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source available"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def hbm_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(args, key=None):
    """dram bytes per launch from the committed ncu capture of this configuration, if any."""
    p = os.path.join(REPO, "profiles", "rollout_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
        return d.get(key or "p%d_%s_g%d_k%d" % (args.players, args.pool, args.games, args.k_steps))
    except Exception:
        return None


def run_b200(args):
    import torch
    import torch.distributed as dist
    from azul_deep_reinforcement_learning_b200.engine import BatchedAzul
    from azul_deep_reinforcement_learning_b200.layout import algorithmic_bytes_per_step

    from azul_deep_reinforcement_learning_b200 import parallel
    rank, world, local = parallel.world()
    if world > 1:
        parallel.init("nccl", local)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pool = 1 if args.pool == "lid" else 0
    G, K = args.games, args.k_steps

    eng = BatchedAzul(G, args.players, pool, 0, seed=args.seed, device=local, game_id_base=parallel.shard(rank, G))
    if args.block:
        eng.set_block_threads(args.block)
    if args.defer:
        eng.set_rollout_defer(args.defer)
    mask = torch.empty((6, G), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        eng.rollout_random(K, mask)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)                    # L2 flush, outside the event bracket
        ev[s][0].record()
        eng.rollout_random(K, mask)
        ev[s][1].record()
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(kernel_ms)
    dev_ms_max = parallel.max_over_ranks(dev_ms, dev)

    # ---- end to end through the host API ------------------------------------------------
    host_state = torch.empty(eng.state.shape, dtype=torch.int32).pin_memory()
    host_state.copy_(eng.state)
    host_mask = torch.empty((6, G), dtype=torch.int32).pin_memory()
    host_cnt = torch.empty(16, dtype=torch.int64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        eng.rollout_random_host(host_state, K, host_mask, host_cnt)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.rollout_random_host(host_state, K, host_mask, host_cnt)
    barrier()
    e2e_s = parallel.max_over_ranks(time.perf_counter() - t0, dev)
    h2d = host_state.numel() * 4
    d2h = host_state.numel() * 4 + host_mask.numel() * 4 + host_cnt.numel() * 8

    # ---- rollout statistics (C2: one small allreduce) -----------------------------------
    cnt = parallel.reduce_counters(eng.counters.clone()).cpu().tolist()

    if rank == 0:
        steps_total = world * G * K * args.steps
        value = steps_total / (dev_ms_max * 1e-3)
        b_alg = algorithmic_bytes_per_step(args.players)
        avg_launch_s = (dev_ms / args.steps) * 1e-3
        achieved = b_alg * G * K / avg_launch_s / 1e9
        peak, peak_src = hbm_peak()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": world * G * K * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": profiled_traffic(args), "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": b_alg, "env_steps_per_launch": G * K,
                         "kernel": "k_rollout_random<%d,%d>" % (args.players, pool),
                         "kernel_ms_avg": dev_ms / args.steps, "kernel_ms_min": min(kernel_ms)},
            "clocks": clocks,
            "wall_s": wall,
            "games_per_sec": (cnt[1] / max(cnt[0], 1)) * value,
            "rollout_counters": {"steps": cnt[0], "games": cnt[1], "rounds": cnt[2], "stuck": cnt[6]},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, n, k, dt = cpu_rollout_rate(args, args.cpu_seconds, threads)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "%d of %d games x %d env steps, C oracle port on %d threads, %.1f s" % (n, G, k, threads, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_policy(args):
    """BASELINE.json configs[3]: self-play, every seat sampled from a random-init ActorCritic (model.py:17-21
    under torch.manual_seed(0)); one bench step = ONE launch of the fused policy kernel = one env step per game."""
    import torch
    from azul_deep_reinforcement_learning_b200 import parallel
    from azul_deep_reinforcement_learning_b200.azulnet.model import ActorCritic
    from azul_deep_reinforcement_learning_b200.engine import BatchedAzul, PackedPolicy, policy_step

    rank, world, local = parallel.world()
    if world > 1:
        parallel.init("nccl", local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pool = 1 if args.pool == "lid" else 0
    G = args.games
    torch.manual_seed(0)
    net = ActorCritic(136, 180)
    eng = BatchedAzul(G, 2, pool, 0, seed=args.seed, device=local, game_id_base=parallel.shard(rank, G))
    packed = PackedPolicy(eng, net)
    lib, h = eng.lib, eng._h
    import ctypes
    out = policy_step(eng, packed, mode=0, apply_step=True, auto_reset=True)      # allocates the output tensors once

    def launch():
        rc = lib.azb_policy_step(h, ctypes.c_void_p(eng.state.data_ptr()), ctypes.c_void_p(packed.buf.data_ptr()), 0, 2,
                                 ctypes.c_void_p(out["action"].data_ptr()), ctypes.c_void_p(out["logp"].data_ptr()),
                                 ctypes.c_void_p(out["value"].data_ptr()), ctypes.c_void_p(out["entropy"].data_ptr()),
                                 ctypes.c_void_p(out["mask"].data_ptr()), ctypes.c_void_p(out["done"].data_ptr()),
                                 ctypes.c_void_p(out["status"].data_ptr()), None, ctypes.c_void_p(eng.counters.data_ptr()), 0,
                                 ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        assert rc == 0

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3) + 40):            # warm-up also spreads the games over all phases of a game
        launch()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        launch()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = parallel.max_over_ranks(ev0.elapsed_time(ev1), dev)

    host_state = torch.empty(eng.state.shape, dtype=torch.int32).pin_memory()
    host_state.copy_(eng.state)
    host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items() if k != "mask"}
    e2e_steps = max(3, min(args.steps, 20))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.state.copy_(host_state, non_blocking=True)
        launch()
        host_state.copy_(eng.state, non_blocking=True)
        for k in host_out:
            host_out[k].copy_(out[k], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    e2e_s = parallel.max_over_ranks(time.perf_counter() - t0, dev)
    cnt = parallel.reduce_counters(eng.counters.clone()).cpu().tolist()
    accuracy = None
    if rank == 0:
        # logits / value of the kernel (fp16 operands, fp32 accumulation) against the untouched fp32
        # ActorCritic forward (model.py:23-41) on the CPU, 4,096 of the mid-game states reached above; outside the timing
        with torch.no_grad():
            chk = policy_step(eng, packed, mode=1, apply_step=False, want_logits=True, want_mask=False)
            k = min(4096, G)
            obs = eng.observe(-1)[:k].cpu()
            ref_l = net.actor_linear2(torch.relu(net.actor_linear1(obs)))
            ref_v = net.forward_critic(obs).squeeze(1)
            got_l, got_v = chk["logits"][:k].cpu(), chk["value"][:k].cpu()
            accuracy = {"states": k, "reference": "fp32 ActorCritic forward on the CPU (torch)",
                        "logits_max_abs_err": float((got_l - ref_l).abs().max()),
                        "logits_mean_abs_err": float((got_l - ref_l).abs().mean()), "logits_max_abs": float(ref_l.abs().max()),
                        "logits_max_err_rel_to_scale": float((got_l - ref_l).abs().max() / ref_l.abs().max()),
                        "value_max_abs_err": float((got_v - ref_v).abs().max())}
    if rank == 0:
        value = world * G * args.steps / (dev_ms * 1e-3)
        flop = 2 * (136 * 360 + 180 * 180 + 180)                                  # 163,080 per decision
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops", 1590.0))
        achieved = flop * G * args.steps / (dev_ms * 1e-3) / 1e12 / world * world / world
        cfg = workload_config(args, world)
        cfg["workload"] = ("%d parallel 2-player self-play games per GPU, every seat sampled from a random-init "
                           "ActorCritic(136,180) by the fused policy kernel (BASELINE.json configs[3])" % G)
        cfg["env_steps_per_game_per_launch"] = 1
        cfg["l2"] = "working set (%.1f MB state + outputs per launch) re-read from L2/HBM every launch; no flush" % (
            eng.state.numel() * 4 / 1e6)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3) + 40, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16 operands, fp32 accumulate (MLP) / u32 (rules)", "data": "synthetic", "config": cfg,
            "e2e": {"value": world * G * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": host_state.numel() * 4,
                    "d2h_bytes_per_step": host_state.numel() * 4 + sum(v.numel() * v.element_size() for v in host_out.values()),
                    "steps": e2e_steps},
            "gpu_launches": args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback",
                         "algorithmic_flop_per_decision": flop, "decisions_per_launch": G, "kernel": "pol::k_policy<%d>" % pool},
            "clocks": clocks,
            "decisions_per_sec": value, "games_per_sec": (cnt[1] / max(cnt[0], 1)) * value, "accuracy": accuracy,
            "rollout_counters": {"steps": cnt[0], "games": cnt[1], "rounds": cnt[2], "stuck": cnt[6]},
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def run_step(args):
    """The single-step entry point the reference's callers use (one Azul.step + next legal mask per call,
    actions supplied by the caller): a genuinely HBM-bound pass -- packed state read + written, 1 action byte in,
    24 mask bytes + done + status out per game -- on a batch larger than L2."""
    import torch
    from azul_deep_reinforcement_learning_b200.engine import BatchedAzul
    from azul_deep_reinforcement_learning_b200.layout import algorithmic_bytes_per_step
    import ctypes
    dev = torch.device("cuda", 0)
    pool = 1 if args.pool == "lid" else 0
    G = args.games
    eng = BatchedAzul(G, args.players, pool, 0, seed=args.seed)
    eng.rollout_random(args.presteps)
    mask = eng.legal_mask()
    done = torch.empty(G, dtype=torch.uint8, device=dev)
    status = torch.empty(G, dtype=torch.uint8, device=dev)
    # one legal action per game, computed once outside the timing: uniformly random among the legal ones (default: the
    # share of games whose round ends in the timed step is then that of random play, ~1 in 10) or the lowest one
    from azul_deep_reinforcement_learning_b200.engine import mask_to_bool
    legal = mask_to_bool(mask)
    if args.step_actions == "random":
        gen = torch.Generator(device=dev).manual_seed(args.seed)
        action = torch.multinomial(legal.float() + 1e-12, 1, generator=gen).squeeze(1).to(torch.uint8)
    else:
        action = legal.to(torch.uint8).argmax(dim=1).to(torch.uint8)
    del legal
    snapshot = eng.state.clone()
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())   # noqa: E731

    def launch():
        assert eng.lib.azb_step(eng._h, p(eng.state), p(action), None, p(mask), None, p(done), p(status), st) == 0

    ms = []
    for i in range(max(args.warmup, 3) + args.steps):
        eng.state.copy_(snapshot)                 # same legal actions every iteration; also evicts nothing we time
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        torch.cuda.synchronize()
        if i >= max(args.warmup, 3):
            ms.append(e0.elapsed_time(e1))
    assert int(status.max()) & 3 == 0
    # MISC word: turn_counter [27:16] moves on new_round, end_of_game [12] when the scored round finished the game
    round_ends = float((((eng.state[3] ^ snapshot[3]) & ((0xFFF << 16) | (1 << 12))) != 0).float().mean())
    b_alg = algorithmic_bytes_per_step(args.players) + 2
    t = statistics.median(ms) * 1e-3
    peak, src = hbm_peak()
    cfg = workload_config(args, 1)
    cfg["workload"] = "%d parallel %d-player games, ONE Azul.step + next legal mask per launch, caller-supplied actions" % (G, args.players)
    cfg["env_steps_per_game_per_launch"] = 1
    cfg["actions"] = "%s legal action per game after %d random-agent steps; %.1f %% of the games end their round (score + refill) in the timed step" % (
        args.step_actions, args.presteps, 100 * round_ends)
    cfg["l2"] = "working set %.0f MB per launch (state %d B + 28 B per game) exceeds the 126 MB L2" % (G * b_alg / 1e6, 4 * eng.W)
    print(json.dumps({
        "metric": METRIC, "value": G / t, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": cfg, "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": G * b_alg / t / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": G * b_alg / t / 1e9 / peak,
                     "traffic": profiled_traffic(args, "step_p%d_%s_g%d_%s" % (args.players, args.pool, G, args.step_actions)),
                     "peak_source": src,
                     "algorithmic_bytes_per_env_step": b_alg, "kernel": "k_step<%d,%d>" % (args.players, pool)}}))


def run_train(args):
    """BASELINE.json configs[4]: the scripts/training.py-equivalent loop -- GPU self-play rollouts of --games
    episodes per rank against the random opponent + one A2C update with a flat NCCL gradient all-reduce per
    bench step."""
    import torch
    from azul_deep_reinforcement_learning_b200 import parallel
    from azul_deep_reinforcement_learning_b200.train import SelfPlayTrainer

    rank, world, local = parallel.world()
    if world > 1:
        parallel.init("nccl", local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    tr = SelfPlayTrainer(args.games, seed=args.seed & 0x7FFFFFFF, device=local, rank=rank, world=world)
    steps_word = lambda: int(tr.runner.engine.state[6].to(torch.int64).sum())   # noqa: E731  env steps executed so far

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        tr.update(tr.rollout())
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env_steps, games, trans, upd_ms, roll_ms, dec_steps = 0, 0, 0.0, [], [], []
    barrier()
    ev0.record()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        s0 = steps_word()
        batch = tr.rollout()
        env_steps += steps_word() - s0           # (rollout resets keep the per-slot step counters running)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        st = tr.update(batch)
        torch.cuda.synchronize()
        upd_ms.append(1e3 * (time.perf_counter() - t1)); roll_ms.append(1e3 * (t1 - t0))
        games += args.games; trans = st["transitions"]; dec_steps.append(int(batch["active"].shape[0]))
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = parallel.max_over_ranks(ev0.elapsed_time(ev1), dev)
    tot = torch.tensor([float(env_steps), float(games)], dtype=torch.float64, device=dev)
    parallel.reduce_counters(tot)
    if rank == 0:
        cfg = workload_config(args, world)
        cfg["workload"] = ("self-play A2C training: %d episodes per GPU per batch vs the random opponent (fused policy kernel "
                           "rollouts), discounted returns, Agent.update loss, Adam, flat NCCL gradient all-reduce "
                           "(BASELINE.json configs[4])" % args.games)
        cfg.pop("env_steps_per_game_per_launch", None)
        cfg["l2"] = "n/a (multi-kernel training step; working set is re-generated every batch)"
        value = float(tot[0]) / (dev_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16-operand rollout MLP / tf32 update (fp32 storage) / u32 rules", "data": "synthetic", "config": cfg,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * 32,
                    "note": "a training batch is end to end by construction: statistics are read back to the host every batch"},
            "gpu_launches": None, "clocks": clocks,
            "games_per_sec": float(tot[1]) / (dev_ms * 1e-3), "agent_decisions_per_batch": trans,
            "rollout_ms": statistics.median(roll_ms), "update_ms": statistics.median(upd_ms),
            "step_ms": [round(a + b, 2) for a, b in zip(roll_ms, upd_ms)], "decisions_per_step": dec_steps,
            "last": {k: tr.history[-1][k] if tr.history else None for k in ()},
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    args = parse_args()
    if not args.games:
        args.games = {"policy": 131072, "train": 1024, "step": 1 << 22}.get(args.mode, 65536)
    if args.mode == "train" and args.impl != "reference":
        return run_train(args)
    if args.mode == "step" and args.impl != "reference":
        return run_step(args)
    if args.mode == "policy" and args.impl != "reference":
        return run_policy(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
