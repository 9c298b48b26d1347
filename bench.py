#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the SMC hot path (BASELINE.json metric) on N B200s.

A "step" (one of --steps K) is ONE COMPLETE FILTER RUN of the workload:
    initialize_particle_filter + (T-1) x [maybe_resample!(ess < N/2) -> particle_filter_step!] + log_ml_estimate
on the 1-D linear-Gaussian state-space model (BASELINE.json configs[2]: Unfold, T=100, 2^24 particles
per GPU, fp64, multinomial resampling, full trace history kept), i.e. N*T particle-steps.

  value      particle-steps/s, device-timed (CUDA events on the library's stream, max over ranks),
             loop enqueued through gsmc_run_steps with no host round trip per step
  e2e        the same run through the Gen-API mirror (initialize_particle_filter /
             maybe_resample_b / particle_filter_step_b / log_ml_estimate): host observation choicemaps
             in, Bool of every maybe_resample! and the log-ML estimate out, allocation included
  roofline   dominant kernel (propagate: proposal + logpdf + ancestor gather + logsumexp partials),
             algorithmic bytes per launch / CUDA-event duration of those launches, vs measured HBM peak
  cpu_baseline   the CPU oracle (C restatement of the reference, 1 thread) on a bounded sample

`--impl reference` times the reference's CPU algorithm (oracle port with all host threads; Julia/Gen.jl
itself cannot run in this environment).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LG = [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0]           # m0, s0, a, b, q, c, r  (SURVEY.md 8(d) cfg 3)
STATE_BYTES = 8                                     # S: one fp64 latent


def make_observations(T):
    """Synthetic observations y_1:T simulated from the model itself (numpy, seed 0); inputs only."""
    m0, s0, a, b, q, c, r = LG
    rng = np.random.default_rng(0)
    x = m0 + s0 * rng.standard_normal()
    ys = []
    for t in range(T):
        if t > 0:
            x = a * x + b + q * rng.standard_normal()
        ys.append(c * x + r * rng.standard_normal())
    return np.array(ys)


def make_bearings_observations(T, sw=0.001, st=0.005, truth=(-0.05, 0.001, 12.0, -0.055)):
    """cfg 5: bearings of a constant-velocity target seen from the origin, simulated with numpy seed 0."""
    rng = np.random.default_rng(0)
    x, vx, y, vy = truth
    obs = []
    for t in range(T):
        if t > 0:
            wx, wy = sw * rng.standard_normal(2)
            x, vx, y, vy = x + vx + 0.5 * wx, vx + wx, y + vy + 0.5 * wy, vy + wy
        obs.append(math.atan2(y, x) + st * rng.standard_normal())
    return np.array(obs)


def workload(args):
    """The configuration being timed. cfg3 (default) is the one BASELINE.json's metric is quoted on; cfg5 is
    BASELINE.json configs[4] (bearings-only, custom proposal, T=200, 2^24 particles per GPU = 2^27 on 8 GPUs)."""
    import gen_b200 as g
    if args.config == "cfg5":
        T = args.T or 200
        model = g.BearingsOnly()
        return {"name": "2D bearings-only tracking (Unfold), T=%d" % T, "model": model, "ys": make_bearings_observations(T), "T": T,
                "proposal": model.custom_proposal(), "proposal_name": "custom (EKF-style) proposal", "S": 32, "kalman": None,
                "kernel": "propagate_kernel<BearingsModel, PROP=1>"}
    T = args.T or 100
    ys = make_observations(T)
    return {"name": "1D linear-Gaussian SSM (Unfold), T=%d" % T, "model": g.LinearGaussianSSM(*LG), "ys": ys, "T": T, "proposal": None,
            "proposal_name": "bootstrap proposal", "S": STATE_BYTES, "kalman": kalman(ys), "kernel": "propagate_kernel<LgssmModel>"}


def claim_stdout():
    """The contract is ONE JSON line on stdout: libraries (NCCL banners, torch warnings) write to fd 1
    too, so fd 1 is pointed at stderr for the whole run and the line goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def emit(saved_fd, line):
    os.write(saved_fd, (json.dumps(line) + "\n").encode())


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (ts, r) in self.rows if t0 - 0.05 <= ts <= t1 + 0.15 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def dist_setup(n_gpus):
    """One process per GPU under torchrun; torch.distributed is plumbing (barrier, id broadcast, max)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return rank, world, local, None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def run_ours(args):
    import gen_b200 as g
    from gen_b200.distributed import Communicator
    rank, world, local, dist = dist_setup(args.gpus)
    n_per = 1 << args.log2n
    N = n_per * world
    wl = workload(args)
    T, ys, model, proposal = wl["T"], wl["ys"], wl["model"], wl["proposal"]
    comm = Communicator(dist, rank, world, device=local) if world > 1 else None

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def new_state():
        return g.ParticleFilterState(model, N, seed=0, dtype=args.dtype, keep_history=not args.no_history,
                                     history_capacity=T, device=local, comm=comm)

    # ---- device-timed value: K complete runs, inputs resident, no host round trip per step ------
    st = new_state()

    def one_run(s):
        s.reset()
        s.init([ys[0]], proposal)
        s.run_steps(ys[1:], N / 2, proposal)
        return s.log_ml_estimate()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        lml = one_run(st)
    launches0 = st.stats()["kernel_launches"]
    barrier()
    t0 = time.time()
    st.synchronize()
    st.timer_start()
    for _ in range(args.steps):
        lml = one_run(st)
    ms = st.timer_stop()
    st.synchronize()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    stats = st.stats()
    launches = stats["kernel_launches"] - launches0
    n_resamples = stats["num_resamples"]
    if dist is not None:
        import torch
        tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    value = N * T * args.steps / (ms * 1e-3)

    # ---- per-kernel CUDA-event times (profiling pass through the per-call API, same workload) ----
    st.set_profiling(True)
    st.reset()
    st.init([ys[0]], proposal)
    remote_rows = None
    for t in range(1, T):
        did = st.maybe_resample(N / 2)
        if did and world > 1 and remote_rows is None and t > T // 2:
            # rows this rank's gather reads over NVLink at one (late) resampling event: ancestors owned by another rank
            anc = st.ancestors()
            remote_rows = int(np.sum((anc < st.first_global) | (anc >= st.first_global + st.num_local)))
            del anc
        st.step([ys[t]], proposal)
    st.log_ml_estimate()
    prof = st.stats()
    st.set_profiling(False)
    st.close()

    # ---- e2e through the Gen-API mirror, host buffers, allocation included --------------------------
    def e2e_run():
        opts = dict(seed=0, dtype=args.dtype, keep_history=not args.no_history, history_capacity=T, device=local, comm=comm)
        if proposal is None:
            state = g.initialize_particle_filter(model, (1,), g.choicemap((model.obs_address(1), float(ys[0]))), N, **opts)
        else:
            state = g.initialize_particle_filter(model, (1,), g.choicemap((model.obs_address(1), float(ys[0]))), proposal, (), N, **opts)
        for Tn in range(2, T + 1):
            g.maybe_resample_b(state)
            g.particle_filter_step_b(state, (Tn,), (g.UnknownChange(),), g.choicemap((model.obs_address(Tn), float(ys[Tn - 1]))), proposal)
        out = g.log_ml_estimate(state)
        state.close()
        return out

    e2e_run()
    barrier()
    e0 = time.perf_counter()
    reps = max(1, min(args.steps, 3))
    for _ in range(reps):
        lml_e2e = e2e_run()
    barrier()
    e_ms = (time.perf_counter() - e0) * 1e3
    if dist is not None:
        import torch
        tmax = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e_ms = float(tmax.item())
    e2e_value = N * T * reps / (e_ms * 1e-3)

    if dist is not None:
        import torch
        rr = torch.tensor([float(remote_rows or 0)], dtype=torch.float64, device="cuda")
        dist.all_reduce(rr, op=dist.ReduceOp.MAX)
        remote_rows = int(rr.item())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_kind = measured_peak()
    S = wl["S"] if args.dtype == "f64" else wl["S"] // 2
    LW = 8 if args.dtype == "f64" else 4
    n_plain, n_gather = prof["n_propagate"], prof["n_propagate_gather"]
    bytes_plain = n_per * (2 * S + 2 * LW)              # read x, lw; write x', lw'   (2S+16)
    bytes_gather = n_per * (2 * S + LW + 4)             # read anc, x[anc]; write x', lw' (lw restarts from 0)
    ms_prop = prof["ms_propagate"] + prof["ms_propagate_gather"]
    alg_bytes = n_plain * bytes_plain + n_gather * bytes_gather
    achieved = alg_bytes / (ms_prop * 1e-3) / 1e9 if ms_prop > 0 else 0.0
    kernel_ms = {k[3:]: prof[k] for k in prof if k.startswith("ms_")}
    total_kernel_ms = sum(kernel_ms.values())
    # whole-run figure with SURVEY 8(d)'s accounting: 2S+16 per particle-step, 2S+36 per resample
    run_bytes = n_per * (T * (2 * S + 2 * LW) + (n_gather) * (2 * S + 36))
    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "%s, N=2^%d particles per GPU, %s, "
                               "multinomial resampling at ESS<N/2, trace history %s" % (wl["name"], args.log2n, wl["proposal_name"], "dropped" if args.no_history else "kept"),
                   "particles_total": N, "time_steps": T, "resamples_per_run": n_resamples,
                   "parallelism": "particles sharded over %d GPU(s)" % world,
                   "arithmetic": "all model arithmetic (transition, logpdf, weights, logsumexp, ESS) in fp64; random numbers: "
                                 "Philox4x32-10, standard normals by a Box-Muller transform evaluated in fp32 on 32-bit words "
                                 "(cuRAND-like resolution) and widened to fp64",
                   "l2": "inputs larger than L2: each step streams >=3 columns of %d MiB (126 MB L2), history columns are never re-read" % (n_per * S >> 20)},
        "log_ml": lml, "log_ml_kalman": wl["kalman"], "log_ml_e2e": lml_e2e,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": 8 * T,
                # per maybe_resample!: 52 B (decision, ESS, log-ML terms + token) stored by the deciding thread into the pinned
                # host mirror; at the end one 432 B copy of the device scalars for log_ml_estimate
                "d2h_bytes_per_step": (T - 1) * 52 + 432,
                "ms_per_step": e_ms / reps,
                "note": "Gen-API mirror; includes allocation of the trace slabs (pooled after the first run); the Bool of every "
                        "maybe_resample! reaches the host through the pinned mirror (device store + token), observations travel as kernel arguments"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": wl["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_kind": peak_kind, "traffic": None,
                     "algorithmic_bytes_per_launch": {"plain": bytes_plain, "gather": bytes_gather},
                     "launches": {"plain": n_plain, "gather": n_gather}, "avg_launch_ms": ms_prop / max(1, n_plain + n_gather)},
        "roofline_whole_run": {"algorithmic_GBps": run_bytes * args.steps / (ms * 1e-3) / 1e9 / 1.0, "frac": run_bytes * args.steps / (ms * 1e-3) / 1e9 / peak,
                               "accounting": "SURVEY 8(d): (2S+16) B per particle-step + (2S+36) B per particle per resample"},
        "kernel_ms_profile_pass": kernel_ms, "kernel_share_propagate": ms_prop / total_kernel_ms if total_kernel_ms else None,
    }
    # second roofline entry: the resampling path (weights/CDF pass + partition + ancestor search) on the events that fired.
    # scan reads lw (8) and writes the CDF (8), search reads the CDF (8) and writes the ancestors (4) = 28 B
    # per particle and event in fp64 (the gather and the zeroed weights are booked with the propagate launch that follows)
    ms_res = prof["ms_scan"] + prof["ms_spacings"] + prof["ms_search"]
    if n_gather and ms_res > 0:
        res_bytes = n_per * (2 * LW + 8 + 4)
        line["roofline_resample"] = {"bound": "hbm", "kernels": "weights_kernel + partition_kernel + search_sorted_kernel",
                                     "achieved": res_bytes * n_gather / (ms_res * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": res_bytes * n_gather / (ms_res * 1e-3) / 1e9 / peak, "events": n_gather,
                                     "algorithmic_bytes_per_event": res_bytes, "avg_event_ms": ms_res / n_gather,
                                     "note": "conditional launches of the steps that did not resample are booked under 'other'"}
    if world > 1:
        line["nvlink"] = {"remote_rows_per_resample_max_rank": remote_rows, "rows_per_rank": n_per,
                          "bytes_per_resample_est": (remote_rows or 0) * (S + 8),
                          "note": "rows of the ancestor gather owned by another rank at one late resampling event (max over ranks); each costs S bytes of "
                                  "state plus the CDF entries of boundary windows over NVLink. Outputs are in ancestor order, so only shard-boundary rows are remote"}
    tr = ncu_traffic()
    if tr and args.dtype == "f64" and args.log2n == 24 and args.config == "cfg3":
        n_l = max(1, n_plain + n_gather)
        line["roofline"]["traffic"] = (n_plain * tr["propagate_plain_bytes"] + n_gather * tr["propagate_gather_bytes"]) / n_l
        line["roofline"]["traffic_source"] = tr.get("source")
        if "roofline_resample" in line and "weights_bytes" in tr:
            line["roofline_resample"]["traffic"] = tr["weights_bytes"] + tr.get("partition_bytes", 0) + tr["search_bytes"]
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(1, args.cpu_log2n, T, ys, args.config)
    emit(args.out_fd, line)
    if dist is not None:
        dist.destroy_process_group()


def kalman(ys):
    """Exact log p(y_1:T) of the linear-Gaussian model (Kalman recursion): the external truth of log_ml."""
    m0, s0, a, b, q, c, r = LG
    m, P, ll = m0, s0 * s0, 0.0
    for t, y in enumerate(ys):
        if t > 0:
            m, P = a * m + b, a * a * P + q * q
        S = c * c * P + r * r
        ll += -0.5 * (y - c * m) ** 2 / S - 0.5 * math.log(2 * math.pi * S)
        K = c * P / S
        m, P = m + K * (y - c * m), (1.0 - K * c) * P
    return ll


def oracle_run(orc, N, T, ys, threads, config="cfg3"):
    from oracle import closed_forms as cf
    from oracle import oracle as O
    fam, params, prop = (O.BEARINGS, list(cf.BEARINGS_PARAMS), 1) if config == "cfg5" else (O.LGSSM, LG, 0)
    pf = orc.particle_filter(fam, params, N, seed=0, keep_history=False, num_threads=threads)
    pf.init([ys[0]], proposal=prop)
    for t in range(1, T):
        pf.maybe_resample()
        pf.step([ys[t]], proposal=prop)
    return pf.log_ml_estimate()


def cpu_baseline(threads, log2n, T, ys, config="cfg3"):
    """The oracle (C port of the reference algorithm) on the host cores, bounded sample."""
    from oracle import oracle as O
    orc = O.Oracle()
    N = 1 << log2n
    t0 = time.perf_counter()
    lml = oracle_run(orc, N, T, ys, threads, config)
    dt = time.perf_counter() - t0
    return {"value": N * T / dt, "unit": "particle-steps/s", "cores": threads, "kind": "port",
            "sample": "same model and loop, N=2^%d particles x T=%d steps, one run (%.1f s); C restatement of Gen.jl's "
                      "algorithm without per-particle trace allocation, so it flatters Gen.jl" % (log2n, T, dt),
            "log_ml": lml}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    T = args.T or (200 if args.config == "cfg5" else 100)
    ys = make_bearings_observations(T) if args.config == "cfg5" else make_observations(T)
    threads = os.cpu_count() or 1
    orc = O.Oracle()
    N = 1 << args.cpu_log2n
    for _ in range(max(1, min(args.warmup, 1))):
        oracle_run(orc, N, T, ys, threads, args.config)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lml = oracle_run(orc, N, T, ys, threads, args.config)
    dt = time.perf_counter() - t0
    value = N * T * args.steps / dt
    sample = "each step = one full filter run on a bounded sample: N=2^%d particles x T=%d (workload is 2^%d per GPU)" % (args.cpu_log2n, T, args.log2n)
    line = {"impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("2D bearings-only tracking (Unfold), T=%d, custom proposal, multinomial resampling at ESS<N/2" if args.config == "cfg5" else
                                    "1D linear-Gaussian SSM (Unfold), T=%d, bootstrap proposal, multinomial resampling at ESS<N/2") % T,
                       "note": "Gen.jl (Julia) cannot run here; this is the C port of its algorithm (oracle/) on all host threads"},
            "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "log_ml": lml, "log_ml_kalman": kalman(ys) if args.config == "cfg3" else None}
    emit(args.out_fd, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=24, help="log2 of particles per GPU")
    ap.add_argument("--T", type=int, default=0, help="time steps (default: 100 for cfg3, 200 for cfg5)")
    ap.add_argument("--config", default="cfg3", choices=["cfg3", "cfg5"], help="cfg3: LG-SSM (the headline); cfg5: bearings-only, custom proposal")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-history", action="store_true")
    ap.add_argument("--cpu-log2n", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.out_fd = claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
