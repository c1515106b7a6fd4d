#!/usr/bin/env python
"""Benchmark of the three-point estimator hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one full estimator call on the workload BASELINE.json quotes the
metric on (config 2): periodic-box bispectrum B_000, full (k1, k2) bin grid
(-> upper triangle, 210 pairs), 10^7 synthetic uniform particles, 512^3 mesh,
PCS assignment (interlacing requested -> switched off by validate(), as in the
reference), 20 linear bins on [0.005, 0.205] h/Mpc.

Own arm (default):
  value   time-to-solution per step in seconds with the catalogue already
          resident in HBM (device pointers into the C-ABI entry point),
          CUDA events on the estimator's stream, max over ranks;
  e2e     the same call fed from PINNED HOST arrays, so every step pays the
          host->device copy of the catalogue and the device->host copy of the
          result vector inside the timed region;
  roofline  the dominant hand-written kernel (particle-to-mesh assignment)
          timed alone with CUDA events against the measured HBM copy bandwidth;
  cpu_baseline  the reference's own C++ (oracle/_ref) on this host's cores, on
          a bounded sample (see `sample`), extrapolated to the full pair count;
  parity_check  the data-vector entries of the bin pairs that the cpu_baseline
          leg computed with the reference's code, compared with the entries the
          GPU produced in the timed region (tolerance 1e-8, BASELINE.json);
  result  SHA-256 of the result vector of one extra step in deterministic mode
          (bit-identical for every N) and head/tail entries of the timed result.
Multi-GPU (torchrun, one rank per GPU): the mesh is replicated, the bin pairs
are dealt to the ranks in compact blocks and one small NCCL all-reduce sums the
result vector; the problem size is fixed, so scaling is "strong".

Reference arm (--impl reference): the reference's CPU implementation
(oracle/_ref) timed on the host cores with every host thread
(TRV_REF_THREADS overrides; torchrun's OMP_NUM_THREADS=1 is not inherited).
Each step is one MEASURED bin-pair unit of the reference's loop (`ms_per_step`
is that measured time); `value` = measured setup + 210 x mean(step) is the
full-job figure and is labelled extrapolated in `sample`.  When the time budget
(TRV_REF_BUDGET_S, default 240 s) allows, one stock reference call is also
measured in full -- trv::compute_bispec_in_gpp_box with form = diag (20 pairs,
the configuration of the reference's JOSS paper) -- and reported under
`measured_full_call`, next to the calibration of the FFTW stand-in against
scipy.fft on the same host (`fft_calibration`).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the assignment
# kernel on the C2 workload, from the committed `ncu --set full` capture
# (the zero-fill memset adds one mesh write, 8 n^3 bytes, on top).
NCU_TRAFFIC = {"k_assign_coop<4,false>": 1.497749e9 + 1.132951e9}
NCU_TRAFFIC_SOURCE = "profiles/r01d_ncu_full_k_assign_coop.csv (kernel only; + 1.07e9 B memset)"

WORKLOADS = {
    "C2": dict(
        name="box B_000 triu, 1e7 uniform particles, 512^3, PCS, 20 lin bins [0.005,0.205]",
        n=10**7, L=1000., ngrid=512, assignment="pcs", degrees=(0, 0, 0), form="full",
        bin_range=(0.005, 0.205), num_bins=20, seed=42,
    ),
    # C2 on the clustered catalogue of SURVEY section 8d (not the default bench line)
    "C2-lognormal": dict(
        name="box B_000 triu, 1e7 lognormal particles (P(k)=2e4 (k/0.05)^-1.5), 512^3, PCS, 20 lin bins",
        n=10**7, L=1000., ngrid=512, assignment="pcs", degrees=(0, 0, 0), form="full",
        bin_range=(0.005, 0.205), num_bins=20, seed=69, catalogue="lognormal",
    ),
    # BASELINE config 5 (the 8-GPU configuration; fits one B200 with 84 GiB): not the
    # default bench line, run with --workload C5 --no-cpu-baseline
    "C5": dict(
        name="box B_000 triu, 1e8 uniform particles, 1024^3, PCS, 40 lin bins [0.005,0.405]",
        n=10**8, L=2000., ngrid=1024, assignment="pcs", degrees=(0, 0, 0), form="full",
        bin_range=(0.005, 0.405), num_bins=40, seed=42,
    ),
    # reduced problem for quick local checks (not a bench line)
    "tiny": dict(
        name="box B_000 triu, 1e5 uniform particles, 64^3, PCS, 6 lin bins",
        n=10**5, L=1000., ngrid=64, assignment="pcs", degrees=(0, 0, 0), form="full",
        bin_range=(0.005, 0.065), num_bins=6, seed=42,
    ),
}


def npairs_of(wl):
    nb = wl["num_bins"]
    return nb * (nb + 1) // 2 if wl["form"] == "full" and wl["degrees"][0] == wl["degrees"][1] \
        else (nb * nb if wl["form"] == "full" else nb)


def lognormal_catalogue(n, L, ngf=256, seed=69):
    """SURVEY section 8d, C2's second catalogue: Gaussian field on 256^3 with
    P(k) = 2e4 (k/0.05)^-1.5 truncated at the Nyquist wavenumber, delta_LN =
    exp(delta_G - sigma^2/2) - 1, Poisson-sampled to ~n points with uniform sub-cell
    jitter, default_rng(69).  Returns exactly n points (resampled if the draw differs)."""
    gen = np.random.default_rng(seed)
    kf = 2 * np.pi / L
    kx = np.fft.fftfreq(ngf, 1. / ngf) * kf
    kz = np.fft.rfftfreq(ngf, 1. / ngf) * kf
    kk = np.sqrt(kx[:, None, None]**2 + kx[None, :, None]**2 + kz[None, None, :]**2)
    pk = np.zeros_like(kk)
    nz = kk > 0
    pk[nz] = 2.e4 * (kk[nz] / 0.05) ** -1.5
    pk[kk > np.pi * ngf / L] = 0.
    white = np.fft.rfftn(gen.normal(size=(ngf, ngf, ngf)))
    dg = np.fft.irfftn(white * np.sqrt(pk / L**3) * ngf**1.5, s=(ngf, ngf, ngf), axes=(0, 1, 2))
    dln = np.exp(dg - dg.var() / 2.)
    cnt = gen.poisson(dln * (n / dln.sum()))
    idx = np.repeat(np.arange(ngf**3), cnt.ravel())
    idx = gen.permutation(idx)                 # catalogue order carries no spatial order
    if idx.size >= n:
        idx = idx[:n]
    else:
        idx = np.concatenate([idx, gen.choice(idx, n - idx.size)])
    i, j, k = np.unravel_index(idx, (ngf, ngf, ngf))
    cell = L / ngf
    return np.ascontiguousarray(np.stack([(i + gen.uniform(size=n)) * cell,
                                          (j + gen.uniform(size=n)) * cell,
                                          (k + gen.uniform(size=n)) * cell]))


def l2_policy(wl):
    """Why no L2 flush is needed between timed iterations, with the workload's own sizes."""
    cat_mb = 24. * wl["n"] / 1.e6
    mesh_gb = 8. * wl["ngrid"] ** 3 / 1.e9
    if cat_mb < 126. or mesh_gb < 0.126:
        return ("inputs fit in the 126 MB L2 (%.0f MB catalogue, %.3f GB meshes): warm-cache timing, "
                "a smoke workload, not a bench line" % (cat_mb, mesh_gb))
    return ("inputs larger than L2 (%.0f MB catalogue, >= %.1f GB meshes); no flush needed"
            % (cat_mb, mesh_gb))


def config_of(args, wl):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": wl["name"], "baseline_config": 5 if args.workload == "C5" else 2,
            "particles": wl["n"], "ngrid": wl["ngrid"], "pairs": npairs_of(wl),
            "l2_policy": l2_policy(wl)}


def host_threads():
    """Threads for the reference's OpenMP code: every host core, whatever
    OMP_NUM_THREADS says (torch.distributed.run exports OMP_NUM_THREADS=1)."""
    env = os.environ.get("TRV_REF_THREADS")
    if env:
        return max(1, int(env))
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def fft_calibration(ref, n=512):
    """The FFTW stand-in of the oracle build against scipy's pocketfft on the same
    host and thread count: one complex n^3 forward transform each."""
    out = {"n": n, "threads": ref.num_threads()}
    try:
        out["shim_s"] = ref.fft_time(n, 2)
    except Exception as exc:   # an older oracle/_ref without the entry point
        out["shim_s"] = None
        out["note"] = f"shim timing unavailable: {exc}"
    try:
        import scipy.fft
        a = np.random.default_rng(0).standard_normal((n, n, n)).astype(np.complex128)
        best = 1.e300
        for _ in range(2):
            t0 = time.perf_counter()
            scipy.fft.fftn(a, workers=out["threads"], overwrite_x=False)
            best = min(best, time.perf_counter() - t0)
        out["scipy_fftn_s"] = best
        if out.get("shim_s"):
            out["shim_over_scipy"] = out["shim_s"] / best
    except Exception as exc:
        out["scipy_fftn_s"] = None
        out["note"] = f"scipy timing unavailable: {exc}"
    return out


def make_catalogue(wl):
    if wl.get("catalogue") == "lognormal":
        return lognormal_catalogue(wl["n"], wl["L"], seed=wl["seed"])
    gen = np.random.default_rng(wl["seed"])
    return gen.uniform(0., wl["L"], size=(3, wl["n"]))


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650., "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    NVML is polled from a background thread (clock + event-reason bitmask only,
    every 20 ms).  A looping ``nvidia-smi -lms`` process was measured to stall
    the CUDA driver for tens of milliseconds per poll on this box -- several
    estimator steps -- so it is only the fallback when NVML cannot be loaded.
    """
    REASONS = {  # nvmlClocksEventReasons bits
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
        0x4: "sw_power_cap",
    }

    def __init__(self, index):
        self.index = index
        self.samples, self.bits = [], 0
        self.thread = self.stop_flag = self.handle = None
        self.sm_max = None
        self.proc = self.file = None

    def start(self):
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = None
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                ent = vis.split(",")[self.index].strip()
                if ent.isdigit():
                    idx = int(ent)
                else:
                    uuid = ent
            self.handle = (pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid
                           else pynvml.nvmlDeviceGetHandleByIndex(idx))
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.stop_flag = threading.Event()

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(
                            self.handle, self.nv.NVML_CLOCK_SM)))
                        self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    except Exception:
                        pass
                    self.stop_flag.wait(0.02)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            self._start_smi()

    def _start_smi(self):
        fields = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                  "clocks_event_reasons.sw_power_cap")
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={fields}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "200"],
                stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "how": "none"}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.sm_max,
                           reasons=sorted(n for b, n in self.REASONS.items() if self.bits & b),
                           samples=len(self.samples), how="nvml thread, 20 ms period")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        rows = [ln.strip().split(", ") for ln in open(self.file.name) if ln.strip()]
        os.unlink(self.file.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)),
                       reasons=sorted(reasons), samples=len(sm), how="nvidia-smi -lms 200")
        return out


# ---------------------------------------------------------------------------
# Own arm
# ---------------------------------------------------------------------------

def run_b200(args):
    import torch
    from triumvirate_b200 import core
    from triumvirate_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ["TRV_GPU_DEVICE"] = str(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        # The one exchange of the path -- the sum of the partial result vectors -- happens
        # inside the estimator call (NCCL behind the C API); torch.distributed ships the id.
        from triumvirate_b200 import dist as tdist
        assert tdist.init_comm() == world

    wl = WORKLOADS[args.workload]
    n, L, ng = wl["n"], wl["L"], wl["ngrid"]
    pos = make_catalogue(wl)
    host = torch.from_numpy(pos).pin_memory()
    dpos = host.to(dev, non_blocking=False)
    torch.cuda.synchronize()
    kw = dict(boxsize=L, ngrid=ng, assignment=wl["assignment"], degrees=wl["degrees"],
              form=wl["form"], bin_range=wl["bin_range"], num_bins=wl["num_bins"],
              norm_factor=1., part_rank=rank, part_count=world)
    dim = npairs_of(wl)

    # Multi-GPU end-to-end: every rank holds the catalogue in pinned host memory, but
    # N simultaneous 240 MB uploads contend for the host's memory and PCIe links
    # (N = 8: 7.4 ms instead of 4.2 ms).  Each rank uploads a 1/N slice and the slices
    # are exchanged over NVLink (one NCCL all-gather of 240 MB).
    share = (n + world - 1) // world
    gathered = torch.empty((3, share * world), dtype=torch.float64, device=dev) if world > 1 else None
    part = torch.empty((3, share), dtype=torch.float64, device=dev) if world > 1 else None

    def upload_sharded(src):
        lo = min(rank * share, n); hi = min(lo + share, n)
        for ax in range(3):   # row slices are contiguous: plain async copies from pinned memory
            part[ax, :hi - lo].copy_(src[ax, lo:hi], non_blocking=True)
        for ax in range(3):
            dist.all_gather_into_tensor(gathered[ax], part[ax])
        torch.cuda.current_stream().synchronize()   # the estimator enqueues on its own stream
        return gathered

    def step(src, on_device, deterministic=False):
        if world > 1 and not on_device:
            src, on_device = upload_sharded(src), True
        out = core.threept_box_arrays("bispec", n, src[0].data_ptr(), src[1].data_ptr(),
                                      src[2].data_ptr(), on_device, deterministic=deterministic,
                                      **kw)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(src, on_device, steps):
        """K steps bracketed by barrier + synchronize; CUDA events on the
        estimator's own stream; returns seconds (max over ranks)."""
        trv = _lib.trv()
        trv.trv_last_stream.restype = C.c_void_p
        sptr = trv.trv_last_stream()
        stream = torch.cuda.ExternalStream(sptr, device=dev) if sptr else torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        per_step = []
        for _ in range(steps):
            ts = time.perf_counter()
            out = step(src, on_device)
            per_step.append(time.perf_counter() - ts)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ev = e0.elapsed_time(e1) * 1.e-3
        if os.environ.get("BENCH_DEBUG"):
            tbl = _lib.trvb()
            tbl.trvb_arena_malloc_count.restype = C.c_longlong
            print(f"[bench debug] on_device={on_device} wall={wall:.4f} ev={ev:.4f} per-step ms="
                  f"{[round(1e3 * t, 2) for t in per_step]} arena cudaMallocs so far="
                  f"{tbl.trvb_arena_malloc_count()}", file=sys.stderr)
        # The estimator synchronises its stream when it returns results, so the
        # event interval and the wall clock agree; keep the larger of the two.
        sec = max(ev, wall if world == 1 else ev)
        if world > 1:
            t = torch.tensor([sec], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec, out

    # warm-up (also builds cuFFT plans, tables and the spline upload)
    for _ in range(max(args.warmup, 3)):
        step(dpos, True)
    tb = _lib.trvb()
    sampler = ClockSampler(local)
    sampler.start()
    tb.trvb_launch_count_reset()
    sec, out = timed(dpos, True, args.steps)
    launches = tb.trvb_launch_count() / args.steps
    fft_execs = tb.trvb_fft_exec_count() / args.steps
    clocks = sampler.stop()

    # end to end from pinned host memory
    for _ in range(2):
        step(host, False)
    sec_e2e, out_e2e = timed(host, False, args.steps)
    assert np.all(np.isfinite(out["bk_raw"].view(np.float64)))

    # One extra step in deterministic mode (outside every timed region): its result
    # vector is bit-identical from run to run and for every N, so its SHA-256 shows
    # that the N = 1, 2, 4, 8 runs computed the same statistics.
    import hashlib
    out_det = step(dpos, True, deterministic=True)
    vec_det = np.concatenate([out_det["bk_raw"].view(np.float64), out_det["bk_shot"].view(np.float64),
                              out_det["k1_eff"], out_det["k2_eff"],
                              out_det["nmodes_1"].astype(np.float64),
                              out_det["nmodes_2"].astype(np.float64)])
    # timed (throughput-mode) result against the deterministic one, per complex entry
    dev_det = float(max(np.max(np.abs(out[k] - out_det[k]) / np.abs(out_det[k]))
                        for k in ("bk_raw", "bk_shot")))

    result = {
        "metric": "bispectrum time-to-solution", "value": sec / args.steps, "unit": "s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1.e3 * sec / args.steps, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args, wl),
        "run": {"parallelism": (
                    (f"{world}gpu: x-slabs of the mesh (assignment, FFTs: all-to-all) and of the "
                     "sub-grid (shell fields, pair products), exchanges inside the C API (NCCL)"
                     if core.dmesh_call_count() > 0 else
                     f"pairs/{world}gpu, replicated mesh, all-reduce inside the C API (NCCL)")
                    if world > 1 else "1gpu"),
                "distributed_mesh_calls": core.dmesh_call_count(),
                "fused_mesh_calls": core.fused_mesh_call_count(),
                "e2e_upload": ("1/N slice per rank from pinned host memory + NCCL all-gather"
                               if world > 1 else "pinned host -> device")},
        "clocks": clocks,
        "e2e": {"value": sec_e2e / args.steps, "unit": "s",
                "h2d_bytes_per_step": int(3 * 8 * n),
                "d2h_bytes_per_step": int(dim * (4 * 8 + 2 * 4 + 4 * 8))},
        "gpu_launches": launches, "cufft_execs": fft_execs,
        "result": {
            "sha256_deterministic_step": hashlib.sha256(vec_det.tobytes()).hexdigest(),
            "timed_vs_deterministic_max_rel": dev_det,
            "bk_raw_first": [float(out["bk_raw"][0].real), float(out["bk_raw"][0].imag)],
            "bk_raw_last": [float(out["bk_raw"][-1].real), float(out["bk_raw"][-1].imag)],
            "bk_shot_first": [float(out["bk_shot"][0].real), float(out["bk_shot"][0].imag)],
            "nmodes_first_last": [int(out["nmodes_1"][0]), int(out["nmodes_2"][-1])],
        },
    }

    # BASELINE config 5 (1024^3, 1e8 particles, 820 pairs -- the configuration north_star
    # names for 8 GPUs) as an extra key of the default line, so that the scaling run records
    # its curve too.  Device-resident catalogue, max over ranks; never fatal for the main line.
    if args.workload == "C2" and not args.no_c5:
        dpos = host = gathered = part = None
        result["c5"] = c5_extra(torch, dist, core, dev, rank, world)
        dpos = torch.from_numpy(pos).to(dev)

    if rank == 0:
        result["roofline"], result["particles_per_s"] = assignment_roofline(torch, dev, dpos, wl)
        if world == 1 and core.fused_mesh_call_count() > 0:
            result["roofline_xpass"] = xpass_roofline(core, step, dpos, wl)
        if world == 1 and not args.no_cpu_baseline:
            base, check = cpu_baseline(wl, pos, out, out_e2e)
            result["cpu_baseline"] = base
            result["parity_check"] = check
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def c5_extra(torch, dist, core, dev, rank, world, steps=3, warmup=2):
    """Time-to-solution of BASELINE config 5 on the GPUs of this run."""
    wl = WORKLOADS["C5"]
    try:
        core.release_contexts()
        torch.cuda.empty_cache()
        free, _ = torch.cuda.mem_get_info(dev)
        if free < 100 * 2**30:
            return {"skipped": f"{free / 2**30:.0f} GiB free HBM, 100 needed"}
        n = wl["n"]
        pos = make_catalogue(wl)
        d = torch.from_numpy(pos).to(dev)
        del pos
        kw = dict(boxsize=wl["L"], ngrid=wl["ngrid"], assignment=wl["assignment"],
                  degrees=wl["degrees"], form=wl["form"], bin_range=wl["bin_range"],
                  num_bins=wl["num_bins"], norm_factor=1., part_rank=rank, part_count=world)

        def step():
            return core.threept_box_arrays("bispec", n, d[0].data_ptr(), d[1].data_ptr(),
                                           d[2].data_ptr(), True, **kw)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(warmup):
            out = step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = step()
        barrier()
        sec = (time.perf_counter() - t0) / steps
        if world > 1:
            t = torch.tensor([sec], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        ok = bool(np.all(np.isfinite(out["bk_raw"].view(np.float64))))
        res = {"workload": wl["name"], "ms_per_step": 1.e3 * sec, "steps": steps, "warmup": warmup,
               "pairs": len(out["bk_raw"]), "finite": ok,
               "bk_raw_first": float(out["bk_raw"][0].real), "nmodes_last": int(out["nmodes_2"][-1]),
               "timing": "wall clock around the calls (each returns its result to the host), max over ranks"}
        del d
        core.release_contexts()
        torch.cuda.empty_cache()
        return res
    except Exception as exc:   # the main line must survive
        try:
            core.release_contexts()
            torch.cuda.empty_cache()
        except Exception:
            pass
        return {"error": f"{type(exc).__name__}: {exc}"[:300]}


def assignment_roofline(torch, dev, dpos, wl):
    """Time the dominant hand-written kernel alone: trvb_assign (zero-fill +
    cell-sorted scatter) of the workload's catalogue onto a REAL mesh.
    Algorithmic bytes per launch = 32 B/particle (24 B position + 8 B weight)
    + one mesh write."""
    from triumvirate_b200 import _lib
    tb = _lib.trvb()
    n, L, ng = wl["n"], wl["L"], wl["ngrid"]
    order = {"ngp": 1, "cic": 2, "tsc": 3, "pcs": 4}[wl["assignment"]]
    ctx = C.c_void_p()
    ngrid = (C.c_int * 3)(ng, ng, ng)
    box = (C.c_double * 3)(L, L, L)

    def chk(st):
        if st != 0:
            raise RuntimeError(tb.trvb_last_error().decode())

    chk(tb.trvb_ctx_create(C.byref(ctx), C.c_int(dev.index or 0), ngrid, box, C.c_int(order)))
    cat = C.c_void_p()
    chk(tb.trvb_cat_create(ctx, C.byref(cat), C.c_longlong(n), C.c_void_p(dpos[0].data_ptr()),
                           C.c_void_p(dpos[1].data_ptr()), C.c_void_p(dpos[2].data_ptr()),
                           None, None, C.c_int(1)))
    mesh = torch.empty(ng * ng * ng, dtype=torch.float64, device=dev)

    class Mesh(C.Structure):
        _fields_ = [("data", C.c_void_p), ("layout", C.c_int), ("k0_add", C.c_double)]

    m = Mesh(mesh.data_ptr(), 0, 0.)
    tb.trvb_ctx_stream.restype = C.c_void_p
    stream = torch.cuda.ExternalStream(tb.trvb_ctx_stream(ctx), device=dev)

    def assign():
        chk(tb.trvb_assign(ctx, cat, C.c_int(0), C.c_int(0), C.c_int(0), C.c_double(1.),
                           C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), m))

    def timed_reps(fn, reps=10):
        for _ in range(3):
            fn()
        tb.trvb_ctx_sync(ctx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        tb.trvb_ctx_sync(ctx)
        return e0.elapsed_time(e1) * 1.e-3 / reps

    def sort_and_assign():
        tb.trvb_cat_invalidate_sort(cat)
        assign()

    t = timed_reps(assign)                 # zero-fill + scatter kernel, sort order cached
    t_sorted = timed_reps(sort_and_assign)  # + counting sort by tile (what an estimator call pays)
    total = float(mesh.sum().item())
    assert abs(total - n) < 1.e-6 * n, "assignment does not conserve the particle count"
    tb.trvb_cat_destroy(cat)
    tb.trvb_ctx_destroy(ctx)
    alg_bytes = 32. * n + 8. * ng**3
    peak, how = measured_peaks()
    achieved = alg_bytes / t / 1.e9
    kname = "k_assign_coop<%d,false>" % order if order >= 3 else "k_assign_scatter<%d,false>" % order
    roof = {"bound": "hbm", "kernel": kname + " (+ zero-fill memset), via trvb_assign",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": NCU_TRAFFIC.get(kname), "traffic_source": NCU_TRAFFIC_SOURCE,
            "peak_source": how,
            "algorithmic_bytes_per_launch": alg_bytes, "launch_seconds": t,
            "with_sort": {"launch_seconds": t_sorted, "achieved": alg_bytes / t_sorted / 1.e9,
                          "frac": alg_bytes / t_sorted / 1.e9 / peak}}
    return roof, n / t_sorted


def xpass_roofline(core, step, dpos, wl):
    """The second hand-written full-grid kernel of the call, k_xpass_fused (forward FFT along
    x + low-|k| modes + shot-noise spectrum + inverse FFT along x, in place on the half
    spectrum), timed with CUDA events on the estimator's stream inside one extra estimator
    call (TRV_XPASS_TRACE=1; outside every timed region).  Algorithmic bytes per launch = one
    read + one write of the half spectrum, 2 x 16 B x n0 n1 (n2/2+1)."""
    from triumvirate_b200 import _lib
    tb = _lib.trvb()
    ng = wl["ngrid"]
    os.environ["TRV_XPASS_TRACE"] = "1"
    try:
        ms = np.zeros((5, 3))
        for r in range(5):
            step(dpos, True)
            buf = (C.c_double * 3)()
            tb.trvb_box_fields_fused_last_ms(buf)
            ms[r] = buf[:]
    finally:
        del os.environ["TRV_XPASS_TRACE"]
    d2z, kern, z2d = np.median(ms, axis=0)
    alg_bytes = 2. * 16. * ng * ng * (ng // 2 + 1)
    peak, how = measured_peaks()
    achieved = alg_bytes / (kern * 1.e-3) / 1.e9
    return {"bound": "hbm", "kernel": "k_xpass_fused<%d> (+ 24 MB low-|k| zero-fill)" % ng,
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (1078631000. + 1044100000.) if ng == 512 else None,
            "traffic_source": ("profiles/r02_ncu_full_k_xpass_fused_async.csv (dram__bytes_read + "
                               "dram__bytes_write of one launch at 512^3)" if ng == 512 else None),
            "peak_source": how, "algorithmic_bytes_per_launch": alg_bytes,
            "launch_seconds": kern * 1.e-3,
            "cufft_2d_d2z_seconds": d2z * 1.e-3, "cufft_2d_z2d_seconds": z2d * 1.e-3}


PARITY_TOL = 1.e-8   # BASELINE.json: relative tolerance against the reference's CPU code


def parity_pairs(nb):
    """Bin pairs the cpu_baseline leg runs through the reference's loop: first and
    last bin, diagonal and off-diagonal."""
    return [(0, 0), (nb - 1, nb - 1), (0, nb - 1), (nb // 3, (2 * nb) // 3)]


def compare_entries(ent, idx, out):
    """Largest deviation of the GPU's entries `out[...][idx]` from the reference's."""
    def rel(a, b):
        return float(np.max(np.abs(a - b) / np.abs(b)))
    return {
        "bk_raw": rel(out["bk_raw"][idx], ent["bk_raw"]),
        "bk_shot": rel(out["bk_shot"][idx], ent["bk_shot"]),
        "k_eff": max(rel(out["k1_eff"][idx], ent["k1_eff"]), rel(out["k2_eff"][idx], ent["k2_eff"])),
        "nmodes_equal": bool(np.array_equal(out["nmodes_1"][idx], ent["nmodes_1"])
                             and np.array_equal(out["nmodes_2"][idx], ent["nmodes_2"])),
    }


def cpu_baseline(wl, pos, out, out_e2e):
    """The reference's own C++ (oracle/_ref) on this host: setup once, then the bin-pair
    units of `parity_pairs`; full job = setup + npairs x mean(unit).  The entries those
    units produce are the oracle's values for the same catalogue at the full mesh: they
    are compared with what the GPU returned in the timed region (`parity_check`)."""
    from oracle import ref
    if not ref.available():
        return ({"value": None, "unit": "s", "kind": "reference", "cores": 0,
                 "sample": "oracle/_ref/libtrv_ref.so missing"},
                {"pairs": 0, "max_rel_err": None, "tol": PARITY_TOL, "passed": False,
                 "note": "oracle/_ref/libtrv_ref.so missing"})
    ref.set_num_threads(host_threads())
    cores = ref.num_threads()
    nb = wl["num_bins"]
    t_setup = ref.bispec_setup(pos, wl["L"], wl["ngrid"], wl["assignment"], wl["bin_range"], nb)
    pairs = parity_pairs(nb)
    ent = ref.bispec_entries(pairs, nb, wl["n"], 1.)
    ref.bispec_teardown()
    units = ent["unit_s"]
    npairs = npairs_of(wl)
    value = t_setup + npairs * float(np.mean(units))
    base = {"value": value, "unit": "s", "cores": cores, "kind": "reference",
            "sample": (f"reference C++ (oracle/_ref, OpenMP on {cores} threads, shim FFT) on the same "
                       f"catalogue and mesh: setup (dn_00, N_L0, G_00, y_lm tables) {t_setup:.2f} s "
                       f"MEASURED once + {len(pairs)} bin-pair units of its loop MEASURED, mean "
                       f"{np.mean(units):.3f} s; value EXTRAPOLATED to {npairs} pairs"),
            "setup_s": t_setup, "pair_unit_s": float(np.mean(units)),
            "measured_s": t_setup + float(np.sum(units))}
    idx = np.array([ref.triu_index(a, b, nb) for a, b in pairs])
    dev = compare_entries(ent, idx, out)
    dev_e2e = compare_entries(ent, idx, out_e2e)
    worst = max(dev["bk_raw"], dev["bk_shot"], dev_e2e["bk_raw"], dev_e2e["bk_shot"])
    check = {"pairs": len(pairs), "bin_pairs": [list(p) for p in pairs],
             "max_rel_err": worst, "tol": PARITY_TOL,
             "passed": bool(worst <= PARITY_TOL and dev["nmodes_equal"] and dev_e2e["nmodes_equal"]
                            and dev["k_eff"] <= 1.e-12 and dev_e2e["k_eff"] <= 1.e-12),
             "device_resident_call": dev, "host_array_call": dev_e2e,
             "against": "reference C++ loop body per bin pair at the full mesh (oracle/_ref: "
                        "S/threept.cpp:1900-1968, 1981-2140)"}
    return base, check


# ---------------------------------------------------------------------------
# Reference arm
# ---------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref
    wl = WORKLOADS[args.workload]
    if not ref.available() and not ref.build():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libtrv_ref.so not built"}))
        return
    t_start = time.perf_counter()
    budget = float(os.environ.get("TRV_REF_BUDGET_S", "240"))
    ref.set_num_threads(host_threads())
    cores = ref.num_threads()
    pos = make_catalogue(wl)
    nb = wl["num_bins"]
    npairs = npairs_of(wl)
    t_setup = ref.bispec_setup(pos, wl["L"], wl["ngrid"], wl["assignment"], wl["bin_range"], nb)
    pairs = [(a, b) for a in range(nb) for b in range(a, nb)]
    want = args.warmup + args.steps
    stride = max(1, len(pairs) // want)
    units, done = [], 0
    for s in range(want):
        a, b = pairs[(s * stride) % len(pairs)]
        _, _, t = ref.bispec_pair(a, b)
        done += 1
        if s >= args.warmup:
            units.append(t)
        # never outrun the lease: stop sampling once the budget is spent (at least one
        # timed unit is always taken)
        if units and time.perf_counter() - t_start > budget:
            break
    ref.bispec_teardown()
    unit = float(np.mean(units))
    value = t_setup + npairs * unit
    steps_done = len(units)
    sample = (f"MEASURED: setup {t_setup:.2f} s (dn_00, N_L0, G_00, y_lm tables) and {steps_done} "
              f"bin-pair units of the reference loop at the full mesh (2 band-limited IFFTs + triple "
              f"product + per-pair shot-noise IFFT and reduction), mean {unit:.3f} s = ms_per_step; "
              f"EXTRAPOLATED: value = setup + {npairs} x mean unit")
    line = {
        "impl": "reference", "metric": "bispectrum time-to-solution", "value": value, "unit": "s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps_done, "warmup": args.warmup,
        "ms_per_step": 1.e3 * unit, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args, wl),
        "cpu_baseline": {"value": value, "unit": "s", "cores": cores, "kind": "reference",
                         "sample": sample, "setup_s": t_setup, "pair_unit_s": unit,
                         "measured_s": t_setup + float(np.sum(units)), "extrapolated": True},
        "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    # One stock call measured in full, when it fits: the `diag` form (nb pairs) costs about
    # setup + nb units.
    est_full = 1.3 * (t_setup + nb * unit)
    left = budget - (time.perf_counter() - t_start)
    if wl["ngrid"] <= 512 and left > est_full:
        full = ref.threept("bispec", "sim", pos, wl["L"], wl["ngrid"], wl["assignment"],
                           wl["degrees"], "diag", wl["bin_range"], nb, 1.)
        line["measured_full_call"] = {
            "what": f"trv::compute_bispec_in_gpp_box, form = diag ({nb} pairs), same catalogue and mesh, "
                    f"entry to result, {cores} threads (the reference's JOSS configuration: 59 s on 32 "
                    f"threads with FFTW, publication/joss/paper.md:269)",
            "seconds": full["elapsed_s"], "pairs": nb,
            "extrapolation_check": {"predicted_s": t_setup + nb * unit,
                                    "ratio": full["elapsed_s"] / (t_setup + nb * unit)}}
    else:
        line["measured_full_call"] = {"skipped": f"estimated {est_full:.0f} s exceeds the remaining "
                                                 f"budget {left:.0f} s (TRV_REF_BUDGET_S)"}
    if budget - (time.perf_counter() - t_start) > 30.:
        line["fft_calibration"] = fft_calibration(ref)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the extra C5 (1024^3) timing")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
