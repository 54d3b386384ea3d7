#!/usr/bin/env python
"""
bench.py - the headline benchmark of BASELINE.json:

    galaxy LM fits/sec (+ model-pixel evals/sec) on N B200s, beside the
    reference algorithm timed on the host's own cores.

Workload (BASELINE.json configs[1]): 10 000 exp-model LM fits per GPU, 48x48
stamps, 3-gaussian ('turb') PSF convolution -> 18 composed gaussians, default
fit_pars (maxfev 4000, ftol = xtol = 1e-5), analytic Jacobian, no prior.
A "step" is one pass of the hot path (ngx_lm_fit) over the whole batch.
Objects shard by index over ranks with no collective: weak scaling, each rank
fits its own 10 000 stamps (seeded by global index).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

prints ONE JSON line (rank 0).  Beside the headline numbers the line carries
"sustained" (the same step back to back for a few seconds) and "configs": the
other BASELINE.json configs (bench_configs.py: single-fit latency, configs[1]
at noise 1.0, 100k EM + admom, 1M bdf fits and the 200k x 5 metacal-style
batch, the last two split over the ranks).

`--impl reference` times the reference's own CPU implementation on all host
cores: the UNMODIFIED ngmix (numba + scipy) through Fitter.go under
multiprocessing when oracle/make_ref.sh has staged it into oracle/_ref
(kind "reference"), else the C restatement in oracle/ (kind "port").
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

_REAL_STDOUT = [None]      # set when stdout is redirected (multi-rank runs)


def emit(line):
    """print the result line on the process's original stdout"""
    if _REAL_STDOUT[0] is None:
        print(line)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT[0], (line + "\n").encode())


NOBJ = 10000
DIMS = (48, 48)
NGAUSS = 18
# algorithmic work per gaussian-pixel evaluation, SURVEY.md section 8(d):
FLOPS_VALUE, FLOPS_DERIV = 28.0, 74.0      # source flops (reference code)
INSTR_VALUE, INSTR_DERIV = 16.0, 30.0      # fused fp64 instructions, ours


# DRAM traffic of one k_lm_fit launch over NCU_LAUNCH_STAMPS stamps, from the
# committed ncu capture (profiles/ncu_lm_fit_r2c.txt)
NCU_DRAM_BYTES_PER_LAUNCH = 377513728 + 114496000
NCU_LAUNCH_STAMPS = 10000


def make_inputs(nobj, seed0, dims=DIMS, noise=0.01):
    """config-2 stamps: parameters on the host (seeded per object), truth
    rendered by the engine itself (exact exp), gaussian noise added"""
    from ngmix_b200 import sims, _lib
    b = sims.make_lm_batch(nobj, seed0=seed0, dims=dims, noise=noise,
                           truth_images=False)
    gm0, st = _lib.gmix_fill("exp", b["truth"])
    assert not st.any()
    gp, _ = _lib.gmix_fill("full", b["psf"].reshape(nobj, -1))
    gm = _lib.gmix_convolve(gm0, gp)
    images, st = _lib.render(gm, b["jac"], dims, fast_exp=False)
    assert not st.any()
    rng = np.random.RandomState(seed0 + 987654321 % (2 ** 31))
    images += rng.normal(0.0, noise, size=images.shape)
    # pixel values as a survey stores them: float32-representable (so the
    # float32 host path below sees exactly the same numbers)
    b["images"] = images.astype(np.float32).astype(np.float64)
    b["weights"] = b["weights"].astype(np.float32).astype(np.float64)
    return b


def make_inputs_host(nobj, seed0, dims=DIMS, noise=0.01):
    """same stamps without a GPU (numpy render) for the reference arm"""
    from ngmix_b200 import sims
    return sims.make_lm_batch(nobj, seed0=seed0, dims=dims, noise=noise)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,"
         "clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip().split(", "))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def time_cpu_port(b, nthreads, target_s=12.0):
    """the oracle port on the host cores, on a bounded sample of the same
    workload; returns (fits/s, nsample, seconds)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    n = b["images"].shape[0]
    # calibrate on a few fits per thread
    ncal = min(n, 4 * nthreads)
    t0 = time.perf_counter()
    oracle.lm_fit_batch("exp", b["images"][:ncal], b["weights"][:ncal],
                        b["jac"][:ncal], b["psf"][:ncal], b["guess"][:ncal],
                        nthreads=nthreads)
    tcal = time.perf_counter() - t0
    rate = ncal / tcal
    ns = int(min(n, max(ncal, rate * target_s)))
    # about target_s seconds of CPU work: the sample is fitted `reps` times
    reps = int(max(1, round(rate * target_s / ns)))
    t0 = time.perf_counter()
    for _ in range(reps):
        res = oracle.lm_fit_batch("exp", b["images"][:ns], b["weights"][:ns],
                                  b["jac"][:ns], b["psf"][:ns],
                                  b["guess"][:ns], nthreads=nthreads)
    dt = time.perf_counter() - t0
    return reps * ns / dt, ns, dt, res, reps


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path
    on the host cores, on a bounded sample of the same workload per step.

    kind "reference" (default when oracle/_ref is staged): the UNMODIFIED
    ngmix (numba + scipy MINPACK) through Fitter.go, one process per core
    with multiprocessing - the north-star's CPU arm (oracle/ref_arm.py).
    kind "port": the C restatement (oracle/), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_arm
    kind = args.kind
    if kind == "auto":
        kind = "reference" if ref_arm.available()[0] else "port"
    if kind == "reference":
        nsample = min(NOBJ, 64 * cores)       # ~0.3 s of CPU work per step
        b = make_inputs_host(nsample, seed0=0)
        r = ref_arm.time_batch(b, kind="lm", model="exp", nproc=cores,
                               steps=args.steps, warmup=max(1, args.warmup),
                               max_seconds=120.0)
        ms, value, nsteps = r["ms_per_step"], r["value"], r["steps"]
        sample = ("%d stamps per step, %d processes (multiprocessing, fork) "
                  "over objects, Fitter.go loop only; %.0f fits/s per core; "
                  "Observation construction %.0f us per object excluded"
                  % (nsample, cores, r["per_core_fits_per_s"],
                     1e6 * r["obs_construction_s_per_object"]))
        impl_note = "ngmix %s (numba + scipy MINPACK), unmodified" \
            % getattr(ref_arm.import_ref(), "__version__", "?")
    else:
        nsample = min(NOBJ, 256 * cores)     # ~0.5 s of CPU work per step
        b = make_inputs_host(nsample, seed0=0)
        times = []
        import oracle
        oracle.build()
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            oracle.lm_fit_batch("exp", b["images"], b["weights"], b["jac"],
                                b["psf"], b["guess"], nthreads=cores)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
            # keep the whole run within a few minutes
            if sum(times) > 120 and len(times) >= 1:
                break
        ms = 1e3 * float(np.mean(times))
        value = nsample / (ms * 1e-3)
        nsteps = len(times)
        sample = "%d stamps per step, %d threads over objects" \
            % (nsample, cores)
        impl_note = "C restatement of the reference algorithm (oracle/)"
    out = {
        "impl": "reference",
        "metric": "galaxy LM fits/sec",
        "value": value, "unit": "fits/s", "n_gpus": args.gpus,
        "steps": nsteps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: exp-model LM fits, 48x48 stamps, "
                               "3-gaussian PSF (18 gaussians), analytic "
                               "Jacobian, no prior, noise 0.01; CPU sample "
                               "of %d stamps per step" % nsample,
                   "implementation": impl_note},
        "cpu_baseline": {"value": value, "unit": "fits/s", "cores": cores,
                         "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "fits/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nobj", type=int, default=NOBJ)
    ap.add_argument("--no-cpu", action="store_true",
                    help="skip the cpu_baseline leg")
    ap.add_argument("--kind", default="auto",
                    choices=["auto", "reference", "port"],
                    help="CPU arm: the staged numba reference or the C port")
    ap.add_argument("--configs", default="all",
                    help="comma list of extra configs (1,2b,3,4,5), 'all' "
                         "or 'none'")
    ap.add_argument("--cfg-steps", type=int, default=2,
                    help="timed steps of each extra config")
    ap.add_argument("--npsf", type=int, default=100000)
    ap.add_argument("--nbdf", type=int, default=1000000)
    ap.add_argument("--nmcal", type=int, default=200000)
    ap.add_argument("--sustain-seconds", type=float, default=3.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from ngmix_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: no CPU fallback")
    torch.cuda.set_device(local)
    numa_node = None
    orig_affinity = os.sched_getaffinity(0)
    if world > 1:
        # one process per GPU: stay on the GPU's own NUMA node, so the pinned
        # buffers of the end-to-end call are local to its PCIe root
        from ngmix_b200 import shard
        numa_node = shard.bind_to_gpu_numa_node(local)
    _lib.init(local)
    if world > 1:
        # keep stdout to the one JSON line without switching NCCL's log off:
        # NCCL writes its NCCL_DEBUG lines (version, INFO with the rank
        # count) to file descriptor 1, so descriptor 1 is pointed at stderr
        # for the rest of the run and the JSON line goes to a private copy of
        # the original stdout
        sys.stdout.flush()
        _REAL_STDOUT[0] = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    nobj = args.nobj
    W = max(args.warmup, 3)

    # roofline denominator: MEASURED_PEAKS.json has no fp64 entry, so the
    # fp64 FMA issue peak is measured here, on this box, before the run
    dfma_per_s, _ = _lib.peak_dfma(60000)

    # ---- inputs: objects [rank*nobj, (rank+1)*nobj)
    b = make_inputs(nobj, seed0=rank * nobj)
    H, Wd = DIMS
    npars = 6

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t

    h_images, h_weights = pinned(b["images"]), pinned(b["weights"])
    h_jac, h_psf, h_guess = pinned(b["jac"]), pinned(b["psf"]), \
        pinned(b["guess"])
    h_res = torch.zeros(nobj * _lib.LM_RESULT_DTYPE.itemsize,
                        dtype=torch.uint8).pin_memory()
    d_images = h_images.to(dev)
    d_weights = h_weights.to(dev)
    d_jac, d_psf, d_guess = h_jac.to(dev), h_psf.to(dev), h_guess.to(dev)
    d_res = torch.zeros(nobj * _lib.LM_RESULT_DTYPE.itemsize,
                        dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        _lib.lm_fit_device("exp", nobj, d_images.data_ptr(),
                           d_weights.data_ptr(), d_jac.data_ptr(), H, Wd,
                           d_psf.data_ptr(), 3, d_guess.data_ptr(),
                           d_res.data_ptr(), stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (inputs in HBM; 369 MB per step >> L2)
    for _ in range(W):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True),
            torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record(stream)
    for k in range(args.steps):
        kev[k][0].record(stream)
        step_device()
        kev[k][1].record(stream)
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = float(np.mean([a.elapsed_time(c) for a, c in kev]))
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps

    res = np.frombuffer(d_res.cpu().numpy().tobytes(),
                        dtype=_lib.LM_RESULT_DTYPE)
    nfev = res["nfev"].astype("f8")
    njev = res["njev"].astype("f8")
    npix = float(H * Wd)
    value_evals = float((nfev * npix).sum())        # model-pixel evals
    deriv_evals = float((njev * npix).sum())
    nflag = int((res["flags"] != 0).sum())

    # ---- end to end through the public host-buffer call: pinned host
    # inputs -> device -> fit -> results back on the host, every step
    e2e_out = np.frombuffer(h_res.numpy(), dtype=_lib.LM_RESULT_DTYPE)

    def step_e2e():
        return _lib.lm_fit("exp", h_images.numpy(), h_weights.numpy(),
                           h_jac.numpy(), h_psf.numpy(), h_guess.numpy(),
                           out=e2e_out)

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r_e2e = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert np.array_equal(r_e2e["flags"], res["flags"])
    # the same call with float32 host stamps (ngx_lm_fit_host_f32): the same
    # pixel values, half the bytes over PCIe
    h_images32 = pinned(b["images"].astype(np.float32))
    h_weights32 = pinned(b["weights"].astype(np.float32))
    assert np.array_equal(h_images32.numpy().astype(np.float64), b["images"])

    def step_e2e32():
        return _lib.lm_fit("exp", h_images32.numpy(), h_weights32.numpy(),
                           h_jac.numpy(), h_psf.numpy(), h_guess.numpy(),
                           out=e2e_out)

    for _ in range(2):
        step_e2e32()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r_e2e32 = step_e2e32()
    torch.cuda.synchronize()
    e2e32_s = time.perf_counter() - t0
    t = torch.tensor([e2e32_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e32_s = float(t.item())
    assert np.array_equal(r_e2e32["pars"], res["pars"], equal_nan=True)
    # the compact survey format: float32 pixels + one weight per stamp (the
    # bench stamps have a uniform weight map and no masked pixels)
    h_ws = pinned(b["weights"][:, 0, 0].copy())
    assert np.array_equal(b["weights"],
                          np.broadcast_to(h_ws.numpy()[:, None, None],
                                          b["weights"].shape))

    def step_e2ec():
        return _lib.lm_fit_compact("exp", h_images32.numpy(), h_ws.numpy(),
                                   h_jac.numpy(), h_psf.numpy(),
                                   h_guess.numpy(), out=e2e_out)

    for _ in range(2):
        step_e2ec()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r_e2ec = step_e2ec()
    torch.cuda.synchronize()
    e2ec_s = time.perf_counter() - t0
    t = torch.tensor([e2ec_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2ec_s = float(t.item())
    assert np.array_equal(r_e2ec["pars"], res["pars"], equal_nan=True)
    # pageable float64 arrays (what a caller who did nothing special has)
    pg_images, pg_weights = b["images"].copy(), b["weights"].copy()

    def step_e2ep():
        return _lib.lm_fit("exp", pg_images, pg_weights, b["jac"], b["psf"],
                           b["guess"])

    step_e2ep()
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(2, args.steps // 3)):
        step_e2ep()
    torch.cuda.synchronize()
    e2ep_s = (time.perf_counter() - t0) / max(2, args.steps // 3)
    t = torch.tensor([e2ep_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2ep_s = float(t.item())
    del pg_images, pg_weights
    h2d = int(h_images.numel() * 8 * 2 + h_jac.numel() * 8
              + h_psf.numel() * 8 + h_guess.numel() * 8)
    d2h = int(nobj * _lib.LM_RESULT_DTYPE.itemsize)

    # ---- the host's host-to-device ceiling with all ranks copying at once
    # (pinned memory, the stamps themselves): what bounds the end-to-end
    # calls when several GPUs share one host
    barrier()
    c0 = torch.cuda.Event(enable_timing=True)
    c1 = torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    for _ in range(4):
        d_images.copy_(h_images, non_blocking=True)
        d_weights.copy_(h_weights, non_blocking=True)
    c1.record(stream)
    barrier()
    cp_ms = c0.elapsed_time(c1)
    t = torch.tensor([cp_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    h2d_gbs = world * 8 * h_images.numel() * 8 / (float(t.item()) * 1e-3) / 1e9

    # ---- the stand-alone fdiff kernel (fill_fdiff + get_loglike fused,
    # ngx_loglike) on the same stamps, at the guess mixtures
    gm0, _ = _lib.gmix_fill("exp", b["guess"])
    gp, _ = _lib.gmix_fill("full", b["psf"].reshape(nobj, -1))
    gmc = _lib.gmix_convolve(gm0, gp)
    _lib.gmix_set_norms(gmc)
    d_gm = torch.from_numpy(
        np.frombuffer(gmc.tobytes(), dtype=np.uint8).copy()).to(dev)
    d_out4 = torch.zeros(nobj * 4, dtype=torch.float64, device=dev)
    d_fd = torch.zeros(nobj * H * Wd, dtype=torch.float64, device=dev)

    def step_fdiff():
        _lib.loglike_device(nobj, NGAUSS, d_gm.data_ptr(),
                            d_images.data_ptr(), d_weights.data_ptr(),
                            d_jac.data_ptr(), H, Wd, d_out4.data_ptr(),
                            d_fdiff=d_fd.data_ptr(),
                            stream=stream.cuda_stream)

    for _ in range(W):
        step_fdiff()
    torch.cuda.synchronize()
    f0 = torch.cuda.Event(enable_timing=True)
    f1 = torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.steps):
        step_fdiff()
    f1.record(stream)
    torch.cuda.synchronize()
    fdiff_ms = f0.elapsed_time(f1) / args.steps

    # ---- sustained rate: the same device-resident step back to back for a
    # few seconds (the K-step timed region above is ~0.1 s)
    sustained = None
    if args.sustain_seconds > 0:
        nrep = int(max(args.steps, args.sustain_seconds * 1e3
                       / max(kernel_ms, 1e-3)))
        s_sampler = ClockSampler(local)
        if rank == 0:
            s_sampler.start()
        barrier()
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(nrep):
            step_device()
        s1.record(stream)
        barrier()
        sus_ms = s0.elapsed_time(s1)
        t = torch.tensor([sus_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sus_ms = float(t.item())
        sc = s_sampler.stop() if rank == 0 else None
        sustained = {"seconds": sus_ms * 1e-3, "steps": nrep,
                     "fits_per_s": world * nobj * nrep / (sus_ms * 1e-3),
                     "clocks": sc}

    # ---- the other BASELINE configs (bench_configs.py)
    import bench_configs
    sel = args.configs.split(",") if args.configs not in ("all", "none") \
        else (["1", "2b", "3", "4", "5"] if args.configs == "all" else [])
    del d_fd, d_out4
    torch.cuda.empty_cache()
    cfg_out = {}
    dd = dist if world > 1 else None
    runners = (("2b", "config2_noise1", bench_configs.config2_noisy),
               ("3", "config3_psf_em_admom", bench_configs.config3_psf),
               ("4", "config4_bdf", bench_configs.config4_bdf),
               ("5", "config5_metacal", bench_configs.config5_metacal))
    for key, name, fn in runners:
        if key not in sel:
            continue
        try:
            cfg_out[name] = fn(args, torch, _lib, dev, rank, world, dd,
                               dfma_per_s)
        except Exception as e:       # a failed extra never loses the headline
            if world > 1:
                raise
            cfg_out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        torch.cuda.empty_cache()
    if "1" in sel and rank == 0:
        try:
            cfg_out["config1_single_fit"] = bench_configs.config1_latency(
                args, _lib)
        except Exception as e:
            cfg_out["config1_single_fit"] = {
                "error": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fits_per_s = world * nobj / (ms_per_step * 1e-3)
    # roofline of the dominant (only) kernel, k_lm_fit: fp64 FMA pipe
    flops = NGAUSS * (value_evals * FLOPS_VALUE + deriv_evals * FLOPS_DERIV)
    instr = NGAUSS * (value_evals * INSTR_VALUE + deriv_evals * INSTR_DERIV)
    peak_flops = 2.0 * dfma_per_s
    achieved = flops / (kernel_ms * 1e-3)
    alg_bytes = nobj * (2 * npix * 8 + 64 + 3 * 48 + 48
                        + _lib.LM_RESULT_DTYPE.itemsize)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    out = {
        "metric": "galaxy LM fits/sec",
        "value": fits_per_s, "unit": "fits/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "configs[1]: %d exp-model LM fits per GPU, 48x48 "
                        "stamps, 3-gaussian PSF (18 gaussians), analytic "
                        "Jacobian, no prior, noise 0.01" % nobj,
            "timing": "inputs resident in HBM; %.0f MB of stamps per step, "
                      "larger than L2" % (alg_bytes / 1e6),
            "parallelism": "objects sharded by index over %d rank(s), no "
                           "collective" % world,
            "numa_node_rank0": numa_node,
        },
        "model_pixel_evals_per_s": world * (value_evals + deriv_evals)
        / (ms_per_step * 1e-3),
        "gaussian_pixel_evals_per_s": world * NGAUSS
        * (value_evals + deriv_evals) / (ms_per_step * 1e-3),
        "fit_stats": {"median_nfev": float(np.median(nfev)),
                      "median_njev": float(np.median(njev)),
                      "nonzero_flags": nflag},
        # the end-to-end headline: the public call with the stamps in the
        # dtype survey pixels have (float32 images and weight maps, pinned
        # host memory); results are bit-identical to the float64 call
        "e2e": {
            "value": world * nobj * args.steps / e2e32_s, "unit": "fits/s",
            "h2d_bytes_per_step": int(h2d - h_images.numel() * 8),
            "d2h_bytes_per_step": d2h,
            "api": "ngmix_b200._lib.lm_fit on float32 host stamps -> "
                   "ngx_lm_fit_host_f32 (pinned host buffers; one "
                   "persistent launch consumes the stamps as the copy "
                   "engine lands them and widens them while staging: "
                   "bit-identical to the float64 call)"},
        "e2e_f64_stamps": {
            "value": world * nobj * args.steps / e2e_s,
            "unit": "fits/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h,
            "api": "ngx_lm_fit_host: the same stamps held as float64 on the "
                   "host (the reference's in-memory dtype): PCIe-bound"},
        "e2e_compact": {
            "value": world * nobj * args.steps / e2ec_s, "unit": "fits/s",
            "h2d_bytes_per_step": int(h2d - h_images.numel() * 12
                                      + nobj * 8),
            "d2h_bytes_per_step": d2h,
            "api": "ngx_lm_fit_host_compact (fit_batch with float32 pixels "
                   "and one weight per stamp): 4 bytes per pixel over PCIe, "
                   "bit-identical results"},
        "host_h2d_ceiling": {
            "aggregate_gbs": h2d_gbs,
            "note": "pinned host-to-device copies of the stamps by all %d "
                    "rank(s) at once; fits/s bound = this / bytes per fit "
                    "(37.1 KB float64, 18.7 KB float32, 9.4 KB compact)"
                    % world,
            "e2e_bound_f64_fits_per_s": h2d_gbs * 1e9 / (h2d / nobj),
            "e2e_bound_f32_fits_per_s": h2d_gbs * 1e9
            / ((h2d - h_images.numel() * 8) / nobj)},
        "e2e_pageable": {
            "value": world * nobj / e2ep_s, "unit": "fits/s",
            "api": "the float64 call on ordinary (pageable) numpy arrays: "
                   "staged into pinned slots by a pool of host threads; "
                   "ngmix_b200.pinned_empty / pinned() give the pinned rate"},
        "gpu_launches": args.steps,   # k_lm_fit launches in the timed region
        "roofline": {
            "bound": "fp64", "achieved": achieved / 1e9,
            "peak": peak_flops / 1e9, "unit": "GFLOP/s",
            "frac": achieved / peak_flops,
            "peak_source": "measured in this run by ngx_peak_dfma (dense "
                           "DFMA issue, x2 flops); MEASURED_PEAKS.json has "
                           "no fp64 entry",
            "algorithmic": "28 flops per value and 74 per derivative "
                           "gaussian-pixel evaluation (SURVEY.md 8d) x 18 "
                           "gaussians x 2304 px x (nfev + njev) passes",
            "frac_fp64_instr": instr / (kernel_ms * 1e-3) / dfma_per_s,
            "kernel": "k_lm_fit", "kernel_ms": kernel_ms,
            "hbm_achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9,
            "hbm_frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one k_lm_fit
            # launch of this workload (10 000 stamps per rank) from the
            # ncu --set full capture in profiles/ncu_lm_fit_r2c.txt:
            # 372.9 MB read + 58.0 MB written; algorithmic 384 MB + records
            "traffic": (NCU_DRAM_BYTES_PER_LAUNCH
                        if nobj == NCU_LAUNCH_STAMPS else None),
            "traffic_source": "ncu capture profiles/ncu_lm_fit_r2c.txt "
                              "(not measured in this run)",
        },
        "fdiff_kernel": {
            "kernel": "k_loglike_tiled (ngx_loglike: fill_fdiff + "
                      "get_loglike + get_model_s2n_sum in one pass)",
            "ms": fdiff_ms,
            "model_pixel_evals_per_s": nobj * npix / (fdiff_ms * 1e-3),
            "gaussian_pixel_evals_per_s": nobj * npix * NGAUSS
            / (fdiff_ms * 1e-3),
            "frac_fp64_flops": nobj * npix * NGAUSS * FLOPS_VALUE
            / (fdiff_ms * 1e-3) / peak_flops,
            "frac_fp64_instr": nobj * npix * NGAUSS * INSTR_VALUE
            / (fdiff_ms * 1e-3) / dfma_per_s,
        },
        "clocks": clocks,
        "sustained": sustained,
        "configs": cfg_out,
    }
    if not args.no_cpu:
        os.sched_setaffinity(0, orig_affinity)    # all host cores again
        cores = os.cpu_count()
        cpu_rate, ns, dt, ores, reps = time_cpu_port(b, cores)
        port = {
            "value": cpu_rate, "unit": "fits/s", "cores": cores,
            "kind": "port",
            "sample": "first %d stamps of the same batch fitted %d times, "
                      "%d threads over objects, %.1f s"
                      % (ns, reps, cores, dt)}
        out["cpu_baseline"] = port
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_arm
        if args.kind != "port" and ref_arm.available()[0]:
            # the north-star arm: unmodified ngmix (numba), one process per
            # core, on the first stamps of the same batch
            nref = min(nobj, 128 * cores)
            bs = {k: b[k][:nref] for k in ("images", "weights", "jac", "psf",
                                           "guess")}
            r = ref_arm.time_batch(bs, kind="lm", model="exp", nproc=cores,
                                   steps=3, warmup=1, max_seconds=60.0)
            out["cpu_baseline"] = {
                "value": r["value"], "unit": "fits/s", "cores": cores,
                "kind": "reference",
                "sample": "first %d stamps of the same batch, %d steps, %d "
                          "processes (multiprocessing) over objects, "
                          "Fitter.go loop only; %.0f fits/s per core"
                          % (nref, r["steps"], cores,
                             r["per_core_fits_per_s"])}
            out["cpu_port"] = port
            out["cpu_port"]["port_over_reference"] = cpu_rate / r["value"]
            rr = r["results"]
            out["parity_vs_reference"] = {
                "n": int(nref),
                "flags_equal": int((rr["flags"] == res["flags"][:nref]).sum()),
                "nfev_equal": int((rr["nfev"] == res["nfev"][:nref]).sum()),
                "max_rel_dev_T": float(np.max(
                    np.abs(res["pars"][:nref, 4] - rr["pars"][:, 4])
                    / np.abs(rr["pars"][:, 4]))),
                "max_rel_dev_flux": float(np.max(
                    np.abs(res["pars"][:nref, 5] - rr["pars"][:, 5])
                    / np.abs(rr["pars"][:, 5]))),
                "max_abs_dev_g": float(np.max(
                    np.abs(res["pars"][:nref, 2:4] - rr["pars"][:, 2:4])))}
        # parity on the sample, reported beside the throughput
        out["parity_on_sample"] = bench_configs.lm_parity(res, ores, 6)
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
