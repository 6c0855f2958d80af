#!/usr/bin/env python
"""bench.py -- ABC particles simulated + scored per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step is one pass of the hot path over one batch of synthetic prior draws: for each of the 5 models,
B particles are drawn (Philox, keyed by the global particle index), simulated with the SSA, reduced to the
53 summary statistics, scored against the 3419 real genes and eps-accepted.  Workload = BASELINE configs[1]
("all 5 models on 1xB200"), streamed in batches; N > 1 shards the particle range over ranks (weak scaling)
and gathers the per-gene acceptance counts and accepted tuples with NCCL.

`value`  : whole-job particles/s with inputs resident in HBM (abc_*_dev entry points, CUDA events).
`e2e`    : the same through the host-buffer C ABI a Julia host calls (abc_simulate_score = abc_simulate + abc_score with the
           error matrix and the accepted lists copied back), wall clock around synchronous calls.
`--impl reference` times the reference's own CPU algorithm (moment ODEs -> 53 statistics -> errors ->
acceptance; oracle port, the Julia/Sundials original cannot run here) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden")

METRIC = "abc_particles_simulated_and_scored_per_s"
UNIT = "particles/s"
SEED = 20240229
EPS = 4.8
# nominal lane-instructions per draw: a six-channel direct-method event and a
# telegraph draw of the hybrid modes (1/4 Philox4x32-10 block 11, uniform -> exponential variate 3, time advance and
# boundary test 2, antiderivative F 5, accumulate and flip the state 3)
NOMINAL_INSTR_PER_EVENT = 64.0
NOMINAL_INSTR_PER_TELEGRAPH_DRAW = 24.0
SSA_MODE = 2                          # ssa_hybrid_burnin of the headline run (library default)
ALG_BYTES_PER_PARTICLE_SCORE = 27776  # 424 B read + G*8 B written, G = 3419
# ncu --set full of abc_tele_kernel, 4096 prior particles of one model (profiles/r2_ssa_m{1,3,5}_ncu_summary.csv): share of issue
# slots used x active threads per instruction / 32 = executed lane-instructions over the lane-issue peak
NCU_SSA = {"src": "profiles/r2_ssa_m{1,3,5}_ncu_summary.csv (ncu --set full --clock-control none of abc_tele_kernel, 4096 prior "
                  "particles of one model, start-time rule)",
           "m1": {"issue_active": 0.662, "threads_per_inst": 29.93, "executed_lane_frac": 0.662 * 29.93 / 32,
                  "warp_inst_per_warp_draw": 25.9},
           "m3": {"issue_active": 0.692, "threads_per_inst": 27.93, "executed_lane_frac": 0.692 * 27.93 / 32},
           "m5": {"issue_active": 0.657, "threads_per_inst": 26.34, "executed_lane_frac": 0.657 * 26.34 / 32},
           "pipes_m1_pct_of_own_peak": {"alu": 42.3, "fma": 36.6, "xu": 31.2}, "dram_bytes_per_launch": 11.0e6}
SCORE_TRAFFIC_131070 = 4.04e9         # measured dram bytes (read + write) of one 131070-particle scoring call, see profiles/


def load_inputs():
    betas = np.load(os.path.join(GOLD, "ref_betas.npy"))
    z = np.load(os.path.join(GOLD, "ref_summary_stats.npz"))
    return betas, z["d"], z["se"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return {"hbm_gbs": j.get("hbm_gbs", 6650.0), "sm_max_mhz": j.get("sm_max_mhz", 1965.0), "src": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  One sampler (rank 0)
    watches all GPUs of the job so that the ranks do not compete with N nvidia-smi processes for the host cores."""

    def __init__(self, indices):
        super().__init__(daemon=True)
        self.indices, self.rows, self.stop_flag = list(indices), [], False

    def run(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        ids = ",".join(str(i) for i in self.indices)
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={ids}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                for line in out.splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            time.sleep(0.25)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        per_gpu = {}
        for r in self.rows:
            try:
                per_gpu.setdefault(r[0], []).append(float(r[1]))
            except (ValueError, IndexError):
                pass
        med = {k: sorted(v)[len(v) // 2] for k, v in per_gpu.items()}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 4 + k and r[4 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": min(med.values()) if med else None, "sm_max_mhz": float(self.rows[0][2]), "reasons": reasons,
                "samples": len(self.rows), "gpus": len(med)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(n_particles, threads, rtol=1e-3, atol=1e-6, seed=SEED):
    """The reference's CPU path on a bounded sample: prior draw -> moment ODEs (run_sim, CVODE-like
    tolerances) -> 53 statistics -> errors vs all genes -> acceptance.  Oracle port, `threads` host threads
    (the C calls release the GIL)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    from abc_inference_transcription_b200 import n_params, split_betas
    betas, d, se = load_inputs()
    od = oracle.make_design(iv_index=1, downsampling=True, betas=split_betas(betas), rtol=rtol, atol=atol)
    oracle.lib()
    jobs = [(1 + (i % 5), i) for i in range(n_particles)]

    def one(job):
        m, i = job
        th = oracle.prior(m, i, seed, n_params(m))
        st, _ = oracle.run_sim(th, m, od)
        err = oracle.compute_trunc_errors(st[None, :], d, se)
        return int((err <= EPS).sum())

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        acc = sum(ex.map(one, jobs))
    dt = time.perf_counter() - t0
    return n_particles / dt, dt, acc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.ref_particles
    vals = []
    for _ in range(args.warmup):
        cpu_reference_sample(max(cores, n // 8), cores)
    t_all = 0.0
    for _ in range(args.steps):
        v, dt, _ = cpu_reference_sample(n, cores)
        vals.append(v)
        t_all += dt
    value = n * args.steps / t_all
    sample = f"{n} prior particles/step over the 5 models (moment-ODE port rtol 1e-3/atol 1e-6 + score vs 3419 genes)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference algorithm = CVODE moment ODEs (scripts/model.jl:89-96); Julia+Sundials cannot run here, "
                    "timed is the C restatement oracle/abc_oracle.c on all host threads"}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "BASELINE configs[1]: all 5 models, prior draws streamed in batches, SSA + 53 statistics + "
                        "error scoring vs 3419 genes + eps=4.8 acceptance",
            "models": 5, "particles_per_model_per_step_per_gpu": args.batch, "n_cells_per_readout": args.n_cells,
            "n_pre_cycles": args.n_pre,
            "burn_in": "ssa_adaptive_burnin=2: the lineages of a (particle, read-out) start at the latest time for which the "
                       "transcripts born earlier contribute in expectation < 2^-n_pre_cycles of the unlabelled and of the labelled "
                       "Poisson mean at the read-out (exact mean contributions per schedule piece; never more than n_pre_cycles "
                       "cycles).  The line also carries ssa_mode2_cycle_burnin (=1, whole cycles per particle, the round-1 rule) "
                       "and ssa_mode2_fixed_burnin (=0, always n_pre_cycles cycles, SURVEY 8d)",
            "readouts": 55, "genes": 3419, "eps": EPS, "seed": SEED,
            "l2": "flushed between steps (256 MiB write)", "lineages": "independent per (condition, age, cell)",
            "ssa": "Gillespie SSA of the gene switch to the read-out, U and L ~ Poisson given the gene path (exact; "
                   "ssa_hybrid_burnin=2), binomial division and capture-efficiency thinning sampled per cell"}


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from abc_inference_transcription_b200 import AbcEngine, ERR_PARTICLE_MAJOR, n_params, synthetic_design
    from abc_inference_transcription_b200.dist import gather_acceptance, init_library_comm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    betas, d, se = load_inputs()
    G = d.shape[0]
    B = args.batch
    eng = AbcEngine(local)
    eng.set_design(synthetic_design(betas, n_cells=args.n_cells, n_pre_cycles=args.n_pre))
    eng.set_data(d, se)
    init_library_comm(eng, world, rank, dev)          # NCCL communicator owned by the library (abc_comm_init_rank)

    stream = torch.cuda.current_stream().cuda_stream
    th_dev = [torch.empty((B, n_params(m)), dtype=torch.float64, device=dev) for m in range(1, 6)]
    st_dev = torch.empty((B, 53), dtype=torch.float64, device=dev)
    err_dev = torch.empty((B, G), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    counts_dev = torch.zeros(G, dtype=torch.int64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def offset_of(step, m):
        # global particle index: disjoint ranges per (step, rank); the same index in different models is a
        # different Philox stream (the model is part of the counter)
        return (step * world + rank) * B

    sim_ms, score_ms, events, draws = [], [], 0, 0

    pending = []          # all-reduces of the per-gene counts in flight (one per step, NCCL's own stream)

    def device_step(step, timed):
        nonlocal events, draws
        eng.accept_reset()
        for m in range(1, 6):
            eng.simulate_dev(m, B, th_dev[m - 1].data_ptr(), st_dev.data_ptr(), particle_offset=offset_of(step, m),
                             seed=SEED, prior_supplied=False, stream=stream)
            eng.score_dev(st_dev.data_ptr(), B, eps=EPS, particle_offset=offset_of(step, m),
                          err_layout=ERR_PARTICLE_MAJOR, d_err_ptr=err_dev.data_ptr(), stream=stream)
            if timed:
                c = eng.counters()       # synchronises; device time of this model's kernels
                sim_ms.append(c["ms_simulate"]); score_ms.append(c["ms_score"])
                events += c["n_events"]; draws += c["n_draws"]
        if world > 1:
            # (i) of SURVEY 8e, once per batch: the G int64 counts are summed over the ranks asynchronously -- nothing in the
            # next batch depends on them, so a rank whose batch was light does not wait for the heaviest one at every step;
            # all of them are awaited before the closing barrier of the timed region
            cbuf = torch.empty(G, dtype=torch.int64, device=dev)
            eng.counts_dev(cbuf.data_ptr(), stream=stream)
            pending.append((dist.all_reduce(cbuf, async_op=True), cbuf))

    def drain_collectives():
        for work, cbuf in pending:
            work.wait()
        pending.clear()

    # ---- device-resident timing ---------------------------------------------------------------
    for w in range(args.warmup):
        device_step(w, False)
        flush.fill_(w & 0xFF)
    drain_collectives()
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(range(world)) if rank == 0 else None
    if sampler:
        sampler.start()
    # EXACTLY `steps` steps between one barrier + synchronize on either side (contract); the L2 flush between steps is a
    # 256 MiB device fill (40 us at HBM speed) kept inside the region rather than stopping all ranks at every step
    e0, e1, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        device_step(args.warmup + k, True)
        flush.fill_(k & 0xFF)
    eb.record()                      # this rank's own work is done; what follows waits for the slowest rank
    drain_collectives()
    e1.record()
    barrier()
    step_ms = [e0.elapsed_time(e1)]
    busy_by_rank = None
    if world > 1:
        tb = torch.tensor([e0.elapsed_time(eb)], dtype=torch.float64, device=dev)
        allb = [torch.zeros_like(tb) for _ in range(world)]
        dist.all_gather(allb, tb)
        busy_by_rank = [round(float(x.item()) / args.steps, 3) for x in allb]
    if sampler:
        sampler.stop_flag = True
    launches = eng.launch_count() - launches0
    t_dev = sum(step_ms) / 1e3
    t = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dev = float(t.item())
    particles = 5 * B * world * args.steps
    value = particles / t_dev

    # ---- end to end through the host-buffer C ABI ----------------------------------------------
    h2d = d2h = 0
    from abc_inference_transcription_b200 import PinnedArray
    # page-locked host buffers (abc_host_alloc), one set per model: theta, statistics, error matrix (224 MB each)
    err_hosts = [PinnedArray((B, G)) for _ in range(5)]
    stats_hosts = [PinnedArray((B, 53)) for _ in range(5)]
    err_host, stats_host = err_hosts[0], stats_hosts[0]
    theta_host = [PinnedArray((B, n_params(m))) for m in range(1, 6)]
    e2e_mode = {"host_theta": True, "note": None}
    counts_dev_host = np.zeros(G, dtype=np.int64)

    def e2e_step(k):
        nonlocal h2d, d2h
        dbg = os.environ.get("ABC_BENCH_DEBUG")
        ta = time.perf_counter()
        eng.accept_reset()
        for m in range(1, 6):
            if dbg:
                print(f"[e2e {k}] m={m} t={1e3 * (time.perf_counter() - ta):.1f} ms", file=sys.stderr, flush=True)
            off = offset_of(args.warmup + k, m)
            if args.e2e_separate:      # the two reference seams as two blocking calls (wrapper.jl section 2, then 3)
                theta, stats, _ = eng.simulate(m, n_trials=B, particle_offset=off, seed=SEED)
                err, counts, _ = eng.score(stats, eps=EPS, particle_offset=off, err_layout=ERR_PARTICLE_MAJOR, out=err_host.array)
                h2d += stats.nbytes
            elif e2e_mode["host_theta"] and not args.e2e_blocking:
                # the reference's sequence: theta = fix_params(...) on the host side of the boundary (abc_simulation.jl:92),
                # then abc_sim(theta, ...) + scoring as one call per (model, batch); theta travels host -> device from
                # page-locked memory inside the timed region.  The calls are the asynchronous ones: the outputs of model m
                # travel to the host while model m + 1 is simulated; abc_wait below completes the step.
                th_in = eng.fix_params(m, B, particle_offset=off, seed=SEED, out=theta_host[m - 1].array)
                eng.simulate_score_async(m, th_in, stats_hosts[m - 1].array, err_hosts[m - 1].array, prior_supplied=True,
                                         particle_offset=off, seed=SEED, eps=EPS, err_layout=ERR_PARTICLE_MAJOR)
                h2d += th_in.nbytes
                theta, stats, err = th_in, stats_hosts[m - 1].array, err_hosts[m - 1].array
                counts = counts_dev_host
            elif e2e_mode["host_theta"]:
                th_in = eng.fix_params(m, B, particle_offset=off, seed=SEED, out=theta_host[m - 1].array)
                theta, stats, err, counts, _ = eng.simulate_score(m, theta=th_in, particle_offset=off, seed=SEED, eps=EPS,
                                                                  err_layout=ERR_PARTICLE_MAJOR, out=err_host.array,
                                                                  stats_out=stats_host.array)
                h2d += th_in.nbytes
            else:                      # prior drawn inside the call (theta is an output only)
                theta, stats, err, counts, _ = eng.simulate_score(m, n_trials=B, particle_offset=off, seed=SEED, eps=EPS,
                                                                  err_layout=ERR_PARTICLE_MAJOR, out=err_host.array,
                                                                  theta_out=theta_host[m - 1].array, stats_out=stats_host.array)
            d2h += theta.nbytes + stats.nbytes + err.nbytes + counts.nbytes
        if e2e_mode["host_theta"] and not args.e2e_blocking and not args.e2e_separate:
            _, wc = eng.wait()                           # every model's theta / statistics / error matrix is on the host
            if dbg:
                print(f"[e2e {k}] device ms: simulate {wc['ms_simulate']:.1f} score {wc['ms_score']:.1f}", file=sys.stderr, flush=True)
        if dbg:
            print(f"[e2e {k}] before gather t={1e3 * (time.perf_counter() - ta):.1f} ms", file=sys.stderr, flush=True)
        res = gather_acceptance(eng, world, dev)     # library-owned NCCL: counts summed, tuples exchanged by gene range
        d2h += res["bytes_d2h"]
        if dbg:
            print(f"[e2e {k}] done t={1e3 * (time.perf_counter() - ta):.1f} ms", file=sys.stderr, flush=True)

    try:                          # untimed warm-up of the host path (work buffers, page-locked staging, sort buffers)
        e2e_step(-2 if args.warmup >= 1 else 0)
        e2e_step(-1 if args.warmup >= 1 else 0)
    except Exception as ex:       # the host-theta sequence failed: time the call that draws the prior itself, and say so
        e2e_mode["host_theta"], e2e_mode["note"] = False, f"host-side fix_params path failed ({ex}); prior drawn inside abc_simulate_score"
        e2e_step(-1 if args.warmup >= 1 else 0)
    h2d = d2h = 0
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_step(k)
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    e2e_value = particles / t_e2e

    # ---- scoring kernel alone on a large resident batch (its own roofline; inputs = simulated prior statistics) ----
    nb = args.score_particles
    # distinct prior particles: their statistics come from the device moment-ODE path (fast), shuffled
    from abc_inference_transcription_b200 import SIM_ODE
    eng_o = AbcEngine(local)
    eng_o.set_design(synthetic_design(betas, sim_kind=SIM_ODE))
    per = (nb + 4) // 5
    big_stats = torch.empty((5 * per, 53), dtype=torch.float64, device=dev)
    for m in range(1, 6):
        th_tmp = torch.empty((per, 9), dtype=torch.float64, device=dev)
        eng_o.simulate_dev(m, per, th_tmp.data_ptr(), big_stats[(m - 1) * per:].data_ptr(), particle_offset=10**7, seed=SEED,
                           stream=stream)
    torch.cuda.synchronize()
    eng_o.close()
    big_stats = big_stats[torch.randperm(5 * per, device=dev)][:nb].contiguous()
    big_err = torch.empty((nb, G), dtype=torch.float64, device=dev)
    score_big_ms = []
    for it in range(4):
        eng.accept_reset()
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.score_dev(big_stats.data_ptr(), nb, eps=EPS, particle_offset=0, err_layout=ERR_PARTICLE_MAJOR,
                      d_err_ptr=big_err.data_ptr(), stream=stream)
        e1.record()
        torch.cuda.synchronize()
        if it > 0:
            score_big_ms.append(e0.elapsed_time(e1))
    score_big_s = sum(score_big_ms) / len(score_big_ms) / 1e3
    # the same call with the tensor-core filter (TF32 tcgen05 GEMM decides which pairs reach the FP64 stage)
    score_mma_ms = []
    eng.set_option("score_mma_filter", 1)
    try:
        for it in range(4):
            eng.accept_reset()
            flush.fill_(it)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.score_dev(big_stats.data_ptr(), nb, eps=EPS, particle_offset=0, err_layout=ERR_PARTICLE_MAJOR,
                          d_err_ptr=big_err.data_ptr(), stream=stream)
            e1.record()
            torch.cuda.synchronize()
            if it > 0:
                score_mma_ms.append(e0.elapsed_time(e1))
    finally:
        eng.set_option("score_mma_filter", 0)
    score_mma_s = sum(score_mma_ms) / len(score_mma_ms) / 1e3
    # acceptance only (no matrix: what a full prior sweep uses, the matrix of 5e6 x 3419 doubles is 137 GB per model)
    from abc_inference_transcription_b200 import ERR_NONE
    score_acc_ms = []
    for it in range(4):
        eng.accept_reset()
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.score_dev(big_stats.data_ptr(), nb, eps=EPS, particle_offset=0, err_layout=ERR_NONE, d_err_ptr=0, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        if it > 0:
            score_acc_ms.append(e0.elapsed_time(e1))
    score_acc_s = sum(score_acc_ms) / len(score_acc_ms) / 1e3
    del big_err, big_stats
    # ---- the same SSA with all six channels simulated explicitly: from the first cycle (mode 0) and inside the
    # ---- label window only (mode 1), small batch ----
    nfd = min(B, 1024)
    cmp_modes = {}
    for mode in (0, 1):
        eng.set_option("ssa_hybrid_burnin", mode)
        ev_fd = ms_fd = 0.0
        for rep in range(2):
            for m in range(1, 6):
                eng.simulate_dev(m, nfd, th_dev[m - 1].data_ptr(), st_dev.data_ptr(), particle_offset=rep * nfd, seed=SEED,
                                 prior_supplied=False, stream=stream)
                c = eng.counters()
                if rep == 1:
                    ev_fd += c["n_events"]; ms_fd += c["ms_simulate"]
        cmp_modes[mode] = {"particles_per_s": 5 * nfd / (ms_fd / 1e3), "events_per_s": ev_fd / (ms_fd / 1e3),
                           "events_per_particle": ev_fd / (5 * nfd), "particles_per_model": nfd,
                           "frac_of_issue_roofline_nominal64": ev_fd / (ms_fd / 1e3) * NOMINAL_INSTR_PER_EVENT / 1e12 /
                                                               (148 * 128 * peaks()["sm_max_mhz"] * 1e6 / 1e12)}
    eng.set_option("ssa_hybrid_burnin", SSA_MODE)
    full_direct = cmp_modes[0]
    pk_instr = 148 * 128 * peaks()["sm_max_mhz"] * 1e6 / 1e12

    def ssa_variant(adaptive, nvar, theta_fn=None, models=(1, 2, 3, 4, 5), reps=2):
        """the same mode-2 kernel under another burn-in rule / on another parameter box: device-resident simulate + score
        (particle-major error matrix) per model, second repetition timed with CUDA events on the launching stream"""
        eng.set_option("ssa_adaptive_burnin", adaptive)
        dr = ms_sim = 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(reps):
            eng.accept_reset()
            if rep == reps - 1:
                torch.cuda.synchronize(); e0.record()
            for m in models:
                th = th_dev[m - 1][:nvar]
                if theta_fn is not None:
                    th.copy_(theta_fn(m, nvar, rep))
                eng.simulate_dev(m, nvar, th.data_ptr(), st_dev.data_ptr(), particle_offset=(7 + rep) * B, seed=SEED,
                                 prior_supplied=theta_fn is not None, stream=stream)
                eng.score_dev(st_dev.data_ptr(), nvar, eps=EPS, particle_offset=(7 + rep) * B, err_layout=ERR_PARTICLE_MAJOR,
                              d_err_ptr=err_dev.data_ptr(), stream=stream)
                if rep == reps - 1:
                    c = eng.counters()
                    dr += c["n_draws"]; ms_sim += c["ms_simulate"]
            if rep == reps - 1:
                e1.record(); torch.cuda.synchronize()
        eng.set_option("ssa_adaptive_burnin", 2)
        t = e0.elapsed_time(e1) / 1e3
        return {"particles_per_s": len(models) * nvar / t, "particles_per_model": nvar, "models": list(models),
                "draws_per_particle": dr / (len(models) * nvar), "draws_per_s": dr / (ms_sim / 1e3),
                "frac_of_issue_roofline_nominal24": dr / (ms_sim / 1e3) * NOMINAL_INSTR_PER_TELEGRAPH_DRAW / 1e12 / pk_instr,
                "ssa_adaptive_burnin": adaptive}

    nv = min(B, 4096)
    fixed_burnin = ssa_variant(0, min(B, 2048))
    cycle_burnin = ssa_variant(1, nv)

    def corner_theta(m, n, rep):
        # BASELINE configs[4] / SURVEY 8d: kon, koff, alpha ~ U(2, 3), gamma ~ U(1, 2) (log10), lambda from its prior
        P = n_params(m)
        g = torch.Generator(device=dev).manual_seed(1000 * m + rep)
        u = torch.rand((n, P), dtype=torch.float64, device=dev, generator=g)
        lo = torch.full((P,), 2.0, dtype=torch.float64, device=dev)
        ngam = 5 if m == 5 else 1
        lo[P - 1 - ngam:P - 1] = 1.0
        th = lo + u
        th[:, P - 1] = -0.7 + 0.7 * u[:, P - 1]
        return th

    corner = ssa_variant(2, B, theta_fn=corner_theta, models=(4, 5), reps=3)
    corner["workload"] = ("BASELINE configs[4]: burst-size / decay-rate models (m = 4, 5) at the high-rate prior corner, kon, koff, "
                          "alpha ~ 10^U(2,3)/h, gamma ~ 10^U(1,2)/h; simulate + score vs 3419 genes, per GPU")
    corner["cycle_burnin"] = ssa_variant(1, min(B, 1024), theta_fn=corner_theta, models=(4, 5))
    sweep = run_full_sweep(args, eng, dev, world, rank, barrier)
    ode = run_ode_path(args, eng_cls=AbcEngine, betas=betas, d=d, se=se, dev=dev, world=world, rank=rank, local=local,
                       barrier=barrier)
    if rank == 0:
        pk = peaks()
        peak_instr = 148 * 128 * pk["sm_max_mhz"] * 1e6 / 1e12        # T lane-instr/s at max clock
        ssa_s = sum(sim_ms) / 1e3
        ev_per_s = events / ssa_s if ssa_s > 0 else 0.0
        draws_per_s = draws / ssa_s if ssa_s > 0 else 0.0
        achieved = draws_per_s * NOMINAL_INSTR_PER_TELEGRAPH_DRAW / 1e12
        sc_s = sum(score_ms) / 1e3
        sc_gbs = nb * ALG_BYTES_PER_PARTICLE_SCORE / score_big_s / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "ms_busy_per_step_by_rank": busy_by_rank,
                "dtype": "f32 (SSA propensities/times) + u32 counts + f64 (statistics, scoring)", "data": "synthetic",
                "config": workload_config(args),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps,
                        "d2h_bytes_per_step": d2h // args.steps,
                        "path": "abc_fix_params -> host theta (page-locked) -> abc_simulate_score_async per model -> abc_wait -> "
                                "abc_accept_fetch (--e2e-blocking: abc_simulate_score per model), two untimed "
                                "warm-up steps" if e2e_mode["host_theta"] and not args.e2e_separate else
                                ("abc_simulate + abc_score" if args.e2e_separate else e2e_mode["note"])},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "roofline": {"kernel": "abc_tele_kernel", "bound": "issue", "achieved": achieved, "peak": peak_instr,
                             "unit": "Tlane-instr/s", "frac": achieved / peak_instr, "traffic": None,
                             "traffic_note": "not memory bound: ncu dram read + write < 10 MB per launch (profiles/r2_ssa_m1_ncu_summary.csv)",
                             "events_per_s": ev_per_s, "events": int(events), "draws": int(draws),
                             "draws_per_s": draws_per_s,
                             "nominal_instr_per_draw": NOMINAL_INSTR_PER_TELEGRAPH_DRAW,
                             "work_unit": "telegraph draw (switch event or sub-interval boundary)",
                             "peak_src": f"148 SM x 128 lanes x sm_max_mhz ({pk['src']})",
                             "ncu": NCU_SSA,
                             "share_of_step": ssa_s / t_dev if t_dev > 0 else None},
                "roofline_score": {"kernel": "abc_score3_classify_kernel + abc_score3_tile_kernel<2> + abc_score3_exact_kernel<2>",
                                   "particles_per_launch": nb, "ms_per_launch": 1e3 * score_big_s,
                                   "pairs_per_s": nb * G / score_big_s, "bound": "hbm", "achieved": sc_gbs, "peak": pk["hbm_gbs"],
                                   "unit": "GB/s", "frac": sc_gbs / pk["hbm_gbs"],
                                   "traffic": SCORE_TRAFFIC_131070 * nb / 131070.0,
                                   "traffic_src": "ncu --set full dram read+write of the three kernels of one 131070-particle "
                                                  "call (profiles/r1_score3_*_ncu_summary.csv), scaled to this launch size",
                                   "peak_src": pk["src"], "share_of_step": sc_s / t_dev if t_dev > 0 else None,
                                   "tensor_core_filter": {"option": "score_mma_filter = 1 (opt-in, bit-identical results)",
                                                          "kernel": "abc_score_mma_prep_kernel + abc_score_mma_filter_kernel<2> (tcgen05 "
                                                                    "TF32 GEMM, TMEM accumulators) + abc_score_mask_exact_kernel<2>",
                                                          "ms_per_launch": 1e3 * score_mma_s,
                                                          "achieved": nb * ALG_BYTES_PER_PARTICLE_SCORE / score_mma_s / 1e9,
                                                          "frac": nb * ALG_BYTES_PER_PARTICLE_SCORE / score_mma_s / 1e9 / pk["hbm_gbs"]},
                                   "accept_only": {"ms_per_launch": 1e3 * score_acc_s, "particles_per_s": nb / score_acc_s,
                                                   "pairs_per_s": nb * G / score_acc_s,
                                                   "note": "err_layout = ABC_ERR_NONE: fused eps-acceptance, no matrix"}}}
        line["ode_path"] = ode
        line["ssa_mode2_fixed_burnin"] = fixed_burnin
        line["ssa_mode2_cycle_burnin"] = cycle_burnin
        line["corner"] = corner
        line["full_sweep"] = sweep
        line["ssa_full_direct"] = full_direct
        line["ssa_six_channel_window"] = cmp_modes[1]
        line["roofline"]["events_per_particle"] = events / max(1, 5 * B * args.steps)
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            v, dt, _ = cpu_reference_sample(args.ref_particles, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.ref_particles} prior particles over the 5 models, reference algorithm "
                                              f"(moment-ODE port, rtol 1e-3) + scoring, {dt:.1f} s"}
            line["cpu_baseline_ssa"] = cpu_ssa_sample(cores, args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_full_sweep(args, eng, dev, world, rank, barrier):
    """BASELINE configs[2] (scripts/accepted_particles.jl:13: 10^6 rows per model): all 5 models, M particles per model in
    total, sharded over the ranks by contiguous global particle ranges (STRONG scaling: the job is fixed, N varies),
    simulated and scored without materialising the error matrix (ABC_ERR_NONE: fused eps-acceptance), then per model the
    counts are all-reduced and the accepted tuples gathered and ordered per gene (v[sortperm(err[v])]).  Wall clock around
    the whole job, max over ranks."""
    import torch
    import torch.distributed as dist
    from abc_inference_transcription_b200 import ERR_NONE, n_params
    from abc_inference_transcription_b200.dist import gather_acceptance, shard_range
    M = args.sweep_particles
    if M <= 0:
        return None
    stream = torch.cuda.current_stream().cuda_stream
    lo, hi = shard_range(M, rank, world)
    n_chunks = max(1, -(-(hi - lo) // args.sweep_chunk))
    chunk = max(1, -(-(hi - lo) // n_chunks))          # equal launches, no short tail
    th = torch.empty((chunk, 9), dtype=torch.float64, device=dev)
    st = torch.empty((chunk, 53), dtype=torch.float64, device=dev)
    accepted, draws, t_gather = [], 0, 0.0
    barrier()
    t0 = time.perf_counter()
    for m in range(1, 6):
        eng.accept_reset()
        for c0 in range(lo, hi, chunk):
            nb = min(chunk, hi - c0)
            eng.simulate_dev(m, nb, th.data_ptr(), st.data_ptr(), particle_offset=c0, seed=SEED + 1, prior_supplied=False, stream=stream)
            eng.score_dev(st.data_ptr(), nb, eps=EPS, particle_offset=c0, err_layout=ERR_NONE, d_err_ptr=0, stream=stream)
            draws += eng.counters()["n_draws"]
        tg = time.perf_counter()
        res = gather_acceptance(eng, world, dev)
        torch.cuda.synchronize()
        t_gather += time.perf_counter() - tg
        accepted.append(int(res["counts"].sum()))
    barrier()
    t = time.perf_counter() - t0
    tt = torch.tensor([t, t_gather], dtype=torch.float64, device=dev)
    dd = torch.tensor([draws], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(dd)
    t, t_gather = float(tt[0]), float(tt[1])
    return {"workload": "BASELINE configs[2]: full prior sweep, all 5 models, simulate + fused eps-acceptance (no error matrix) + "
                        "per-gene ordered accepted lists; strong scaling over --gpus",
            "particles_per_model": M, "models": 5, "n_gpus": world, "scaling": "strong", "wall_s": t,
            "particles_per_s": 5 * M / t, "gather_and_order_s": t_gather, "accepted_pairs_per_model": accepted,
            "draws_per_particle": float(dd[0]) / (5 * M), "chunk": chunk,
            "extrapolated_wall_s_at_5e6_per_model": t * 5e6 / M}


def run_ode_path(args, eng_cls, betas, d, se, dev, world, rank, local, barrier):
    """Same step with sim_kind = ABC_SIM_ODE: the moment-ODE computation the reference's CPU path performs
    (scripts/model.jl), on the device.  Reported next to the SSA headline, never instead of it."""
    import torch
    import torch.distributed as dist
    from abc_inference_transcription_b200 import ERR_PARTICLE_MAJOR, SIM_ODE, n_params, synthetic_design
    B = args.ode_batch
    G = d.shape[0]
    eng = eng_cls(local)
    eng.set_design(synthetic_design(betas, sim_kind=SIM_ODE))
    eng.set_data(d, se)
    stream = torch.cuda.current_stream().cuda_stream
    th_dev = [torch.empty((B, n_params(m)), dtype=torch.float64, device=dev) for m in range(1, 6)]
    st_dev = torch.empty((B, 53), dtype=torch.float64, device=dev)
    err_dev = torch.empty((B, G), dtype=torch.float64, device=dev)
    steps = max(1, args.steps)
    from abc_inference_transcription_b200 import PinnedArray
    err_host = PinnedArray((B, G))

    def dev_step(k):
        eng.accept_reset()
        for m in range(1, 6):
            off = (k * world + rank) * B
            eng.simulate_dev(m, B, th_dev[m - 1].data_ptr(), st_dev.data_ptr(), particle_offset=off, seed=SEED, stream=stream)
            eng.score_dev(st_dev.data_ptr(), B, eps=EPS, particle_offset=off, err_layout=ERR_PARTICLE_MAJOR,
                          d_err_ptr=err_dev.data_ptr(), stream=stream)

    dev_step(0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        dev_step(1 + k)
    e1.record()
    barrier()
    t_dev = e0.elapsed_time(e1) / 1e3
    ode_steps = eng.counters()["n_ode_steps"]

    def host_step(k):
        eng.accept_reset()
        for m in range(1, 6):
            off = ((1 + k) * world + rank) * B
            theta, stats, _ = eng.simulate(m, n_trials=B, particle_offset=off, seed=SEED)
            err, counts, _ = eng.score(stats, eps=EPS, particle_offset=off, err_layout=ERR_PARTICLE_MAJOR, out=err_host.array)
        eng.accept_fetch()

    host_step(0)                  # untimed warm-up of the host path
    barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        host_step(k)
    barrier()
    t_e2e = time.perf_counter() - t0
    tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    n = 5 * B * world * steps
    eng.close()
    return {"sim_kind": "moment ODEs on device (Radau IIA, rtol 1e-6)", "value": n / float(tt[0]), "unit": UNIT,
            "e2e": n / float(tt[1]), "particles_per_model_per_step_per_gpu": B, "steps": steps,
            "ode_steps_last_launch": int(ode_steps)}


def cpu_ssa_sample(cores, args):
    """same SSA workload on the host cores (oracle SSA port), bounded: a few read-outs of a few particles"""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    from abc_inference_transcription_b200 import n_params, split_betas
    betas, _, _ = load_inputs()
    sd, keep = oracle.make_ssa_design(args.n_cells, args.n_pre, True, split_betas(betas))
    jobs = [(1 + (i % 5), i, (7 * i) % 11, (3 * i) % 5) for i in range(16 * cores)]

    def one(job):
        m, i, c, a = job
        th = oracle.prior(m, i, SEED, n_params(m))
        _, ev = oracle.ssa_readout(th, m, sd, i, SEED, c, a, oracle.MATH_LIBM)
        return ev

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        ev = sum(ex.map(one, jobs))
    dt = time.perf_counter() - t0
    readouts_per_s = len(jobs) / dt
    return {"value": readouts_per_s / 55.0, "unit": UNIT, "cores": cores, "kind": "port",
            "events_per_s": ev / dt,
            "sample": f"{len(jobs)} read-outs ({args.n_cells} cells each) of prior particles, oracle SSA, {dt:.1f} s; "
                      "particles/s = read-outs/s / 55"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8192, help="particles per model per step per GPU")
    ap.add_argument("--n-cells", type=int, default=96)
    ap.add_argument("--n-pre", type=int, default=10)
    ap.add_argument("--ode-batch", type=int, default=32768, help="particles per model per step for the ODE-path line")
    ap.add_argument("--score-particles", type=int, default=131072, help="particles per launch for the scoring-kernel roofline")
    ap.add_argument("--ref-particles", type=int, default=24000, help="particles per bounded CPU sample (~10 s on 16 threads)")
    ap.add_argument("--sweep-particles", type=int, default=1000000, help="full_sweep: particles per model in total (0 = skip)")
    ap.add_argument("--sweep-chunk", type=int, default=65536, help="full_sweep: particles per launch per rank")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-blocking", action="store_true", help="e2e through the blocking abc_simulate_score (one call per model)")
    ap.add_argument("--e2e-separate", action="store_true", help="e2e through abc_simulate + abc_score instead of abc_simulate_score")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
