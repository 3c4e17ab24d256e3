#!/usr/bin/env python
"""Headline benchmark: atom-steps/s of the MTP pair-style compute (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2]

A "step" is one full force evaluation (Pair::compute with eflag=1, vflag=1) over the workload:
configs[1] of BASELINE.json by default (bcc W/Mo, 262,144 atoms, MTP level 16, 2 species, positions =
lattice + 0.05 A jitter, full neighbor list at cutoff 5 A + 2 A skin, ghosts explicit like LAMMPS).
N > 1 (torchrun, one rank per GPU): the global box is the per-GPU box replicated on a brick grid
{2x1x1, 2x2x1, 2x2x2}; every step each rank packs its halo on the device, exchanges ghost positions
and ghost forces with NCCL send/recv and runs the same kernels on its brick (weak scaling).

One JSON line is printed by rank 0 (contract in the task statement): `value` = device-timed whole-job
atom-steps/s with inputs resident in HBM, `e2e` = the same metric through the host-buffer C-ABI call
(mtp_compute_host: H2D of x/type/f/neighbor list, compute, D2H of f and the energy/virial record),
`roofline` for the dominant kernel against the FP64 peak measured in this run, `cpu_baseline` = the
reference's own CPU `mtp` style timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "lammps-mtp-kokkos_b200"))

METRIC = "atom-steps/sec (Matom·step/s) at 1/2/4/8 B200 vs CPU mtp; % of FP64/HBM roofline"
UNIT = "Matom·step/s"


# ----------------------------------------------------------------------------------------------- helpers
def algorithmic_flops_per_atom(pot, n_list, n_cut):
    """flops counted as written in pair_mtp.cpp (add/mul/div/sqrt = 1, fma = 2)."""
    B, R, P = pot.radial_basis_size, pot.radial_funcs_count, pot.max_alpha_index_basic
    K, T, A = pot.K, pot.T, pot.A
    nz = (np.asarray(pot.alpha_index_basic)[:, 1:] != 0).sum(axis=1)
    per_pair = (8 * B + 19) + 4 * (P - 1) + 4 * R * B + float((15 + 5 * nz).sum()) + 6 * K + 6
    return 9.0 * n_list + n_cut * per_pair + 9.0 * T + 2.0 * A


def algorithmic_flops_split(pot, n_list, n_cut):
    """The same count apportioned to the kernels of the pipeline: the gather kernel owns
    gather/cutoff, Chebyshev and the radial contraction; the moment kernel the powers and the m[k] accumulation (6 flop
    per k: nf, val, pw x2, fma); the force kernel the Jacobian part (9 + 5 nz_k) and the contraction with dE/dm
    (6K + 6); the program kernel the contraction tree forward/reverse and the energy."""
    B, R, P = pot.radial_basis_size, pot.radial_funcs_count, pot.max_alpha_index_basic
    K, T, A = pot.K, pot.T, pot.A
    nz = (np.asarray(pot.alpha_index_basic)[:, 1:] != 0).sum(axis=1)
    gather = 9.0 * n_list + n_cut * ((8 * B + 19) + 4 * R * B)
    moments = n_cut * (4 * (P - 1) + 6.0 * K)
    forces = n_cut * (float((9 + 5 * nz).sum()) + 6 * K + 6)
    program = 9.0 * T + 2.0 * A
    return {"gather": gather, "moments": moments, "forces": forces, "program": program}


def executed_flops_split(pot, n_list, n_cut, program_terms):
    """FP64 operations this design actually executes per atom (fma = 2 flop), kernel by kernel -- next to the
    reference-convention count above, which charges the Jacobian the reference forms (pair_mtp.cpp:175-191) and this
    design never does.  Standard basic-moment sets only (degree D0 = P - 1).
      gather+radial: r, rsq per listed neighbor; per in-cutoff pair sqrt, 1/d, u, Chebyshev values + derivatives by
                     recurrence and the two radial contractions (2 x R x B fma);
      moments:       one fma per (pair, basic moment) + one multiply per monomial u^q;
      forces:        Horner gradient of sum_mu f_mu(d) P_mu(u): two fma per (pair, basic moment) for W / W', 3 / 4 / 5 fma
                     per monomial / (a, b) column / a level, ~30 for the epilogue, scatter and virial;
      program:       one fma per term step of the generated kernel (forward T + reverse 2T with squares merged)."""
    B, R, P, K = pot.radial_basis_size, pot.radial_funcs_count, pot.max_alpha_index_basic, pot.K
    d0 = P - 1
    nq, nab, na = (d0 + 1) * (d0 + 2) * (d0 + 3) // 6, (d0 + 1) * (d0 + 2) // 2, d0 + 1
    gather = 8.0 * n_list + n_cut * (2 + 3 + 10 + 9.0 * max(B - 2, 0) + 4.0 * R * B)
    moments = n_cut * (2.0 * K + nq)
    forces = n_cut * (4.0 * K + 6.0 * nq + 8.0 * nab + 10.0 * na + 30.0)
    program = 2.0 * program_terms
    return {"gather": gather, "moments": moments, "forces": forces, "program": program}


def algorithmic_bytes_per_atom(n_list):
    return 4.0 * n_list + 84.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.index = index
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [s for s, p in zip(sm, pw) if p > 300] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def brick_grid(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


LIST_EVERY = 10      # e2e: steps between neighbor-list uploads (re-neighboring cadence; the default skin of 2 A gives far more)


def pinned_like(a):
    import torch
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, pin_memory=True)
    v = t.numpy()
    v[...] = a
    return t, v


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_model():
    """Host CPU model string (so that two rounds' reference arms can be compared)."""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, cfg, pot_path, pot):
    """The reference's own CPU `mtp` (oracle/_ref, unmodified sources) on all host cores; falls back to
    the C restatement (kind "port") only if the prebuilt reference library did not travel."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    from concurrent.futures import ThreadPoolExecutor
    from mtp_b200 import harness

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cells = args.ref_cells
    sysm = harness.make_config(args.config, cells=cells)
    kind = "reference" if os.path.exists(oracle_py.REF_SO) else "port"
    if kind == "port":
        oracle_py.build(ref=False)
    slabs = np.array_split(sysm.ilist, cores)

    def make():
        return oracle_py.ReferenceMTP("mtp", pot_path) if kind == "reference" else oracle_py.OracleMTP(pot)

    workers = [make() for _ in range(cores)]

    def one(k):
        w, il = workers[k], slabs[k]
        if kind == "reference":
            r = w.compute(sysm.x, sysm.type, sysm.nlocal, il, sysm.numneigh, sysm.neigh, sysm.offsets, eflag=1, vflag=1)
        else:
            r = w.compute(sysm.x, sysm.type, il, sysm.numneigh, sysm.neigh, sysm.offsets, eflag=1, vflag=1)
        return r.energy

    times = []
    with ThreadPoolExecutor(max_workers=cores) as ex:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            es = list(ex.map(one, range(cores)))
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sysm.nlocal / (ms * 1e-3) / 1e6
    sample = (f"{sysm.nlocal}-atom sub-box ({cells[0]}x{cells[1]}x{cells[2]} cells) of the same lattice/potential, "
              f"{cores} threads each owning a contiguous slab of ilist (PairMTP::compute, eflag=1 vflag=1), "
              f"g++ -O2 -ffp-contract=off")
    return dict(value=value, ms_per_step=ms, cores=cores, kind=kind, sample=sample, energy=float(sum(es)),
                natoms=sysm.nlocal)


# ----------------------------------------------------------------------------------------------- main
_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # native libraries (NCCL's version banner, ...) write to file descriptor 1: keep it for the JSON line only
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--cells", type=int, nargs=3, default=None, help="override the per-GPU cell counts")
    ap.add_argument("--ref-cells", type=int, nargs=3, default=None, help="CPU-baseline sample box")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="large", choices=["large", "small"])
    ap.add_argument("--device-list", default="auto", choices=["auto", "on", "off"],
                    help="build the neighbor list on the device (mtp_neigh_build) instead of on the host; auto = on above 1.5 M "
                         "atoms per GPU (config 5), where a host build and a host copy of the list make no sense")
    ap.add_argument("--grade-every", type=int, default=0,
                    help="config 4: request per-atom extrapolation grades every K-th step (fix pair semantics); 0 = never")
    ap.add_argument("--halo", default="direct", choices=["direct", "staged"],
                    help="ghost exchange: one 26-direction stage (default) or LAMMPS's three dimension-by-dimension stages")
    ap.add_argument("--md-steps", type=int, default=50,
                    help="informational device-resident NVE run of this many steps after the bench (N = 1 only; 0 = off)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N > 1: do not hide the halo exchange behind the interior atoms (decomp.OverlappedStep + "
                         "mtp_compute_phased); default is to overlap")
    ap.add_argument("--lanes", type=int, default=3, help="internal streams the super-chunks are dealt to (mtp_set_lanes)")
    ap.add_argument("--chunksize", type=int, default=131072,
                    help="pair_style ... chunksize N: README.md:44 of the reference asks the user to tune it (\"sufficient "
                         "parallelism\", \"minimizing the occurrence of a small final chunk\"); 131072 is the tuned value for "
                         "this implementation at config 2 (sweep in DESIGN.md section 6); the reference's decks use 32768")
    args = ap.parse_args()

    from mtp_b200 import almtp, harness
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = harness.CONFIGS[args.config]
    if args.ref_cells is None:
        per_cell = {"sc": 1, "bcc": 2, "fcc": 4, "diamond": 8}[cfg["kind"]]
        n = 2
        while per_cell * (n + 2) ** 3 <= 28000:
            n += 2
        args.ref_cells = (n, n, n)

    tmp = tempfile.mkdtemp(prefix="mtp_bench_")
    pot_path = os.path.join(tmp, f"config{args.config}.almtp")
    pot = almtp.random_potential(cfg["level"], cfg["species"], with_active_set=bool(cfg.get("active_set")))
    if rank == 0 or args.impl == "b200":
        almtp.write_almtp(pot_path if world == 1 else pot_path + f".{rank}", pot)
        if world > 1:
            pot_path = pot_path + f".{rank}"
    pot = almtp.read_almtp(pot_path) if os.path.exists(pot_path) else pot
    cells = tuple(args.cells) if args.cells else cfg["cells"]
    workload = (f"config[{args.config - 1}]: {cfg['name']}; {cells[0]}x{cells[1]}x{cells[2]} {cfg['kind']} cells per GPU, "
                f"jitter 0.05 A, cutoff 5.0 A + skin 2.0 A, random-init MLIP-3 coefficients (seed = level)")

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        r = run_reference(args, cfg, pot_path, pot)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "sample": r["sample"]},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"], "cpu_model": cpu_model()},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from mtp_b200 import api
    from mtp_b200.api import MTPB200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the MTP B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's version / debug lines must not land on stdout
        dist.init_process_group("nccl", device_id=dev)

    variant = api.VARIANT_SMALL if args.variant == "small" else api.VARIANT_LARGE
    mtp = MTPB200(pot_path, selection_state=bool(cfg.get("active_set")), device=local_rank)
    mtp.set_chunksize(args.chunksize)
    mtp.set_lanes(args.lanes)

    # one brick per rank on the grid {1, 2x1x1, 2x2x1, 2x2x2}; at N = 1 all six swaps are periodic self-images
    from mtp_b200 import decomp
    per_cell = {"sc": 1, "bcc": 2, "fcc": 4, "diamond": 8}[cfg["kind"]]
    devlist_mode = args.device_list == "on" or (args.device_list == "auto" and per_cell * cells[0] * cells[1] * cells[2] > 1500000)
    sysm, halo = decomp.make_rank_system(args.config, cells, brick_grid(world), rank, dev, mtp.lib,
                                         direct=args.halo == "direct", with_list=not devlist_mode)
    nlocal, nall = sysm.nlocal, sysm.nall

    # device-resident inputs
    t_x = torch.from_numpy(sysm.x).to(dev)
    t_type = torch.from_numpy(sysm.type).to(dev)
    t_ilist = torch.from_numpy(sysm.ilist).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    if devlist_mode:
        # the list is built on the device from the uploaded positions (ghost rows included); row-major table
        nn_d, t_neigh, max_nn = mtp.neigh_build(t_x, nlocal, sysm.rlist, stream=stream)
        t_nn = torch.zeros(nall, dtype=torch.int32, device=dev)
        t_nn[:nlocal] = nn_d
        t_off, list_stride = None, int(t_neigh.shape[1])
        n_list = float(nn_d.double().mean().item())
        ns = min(nlocal, 65536)      # in-cutoff neighbors per atom from the first 65536 rows
        rows = t_neigh[:ns].long().clamp_(0, nall - 1)
        d2 = ((t_x[rows] - t_x[:ns, None, :]) ** 2).sum(dim=2)
        live = torch.arange(list_stride, device=dev)[None, :] < nn_d[:ns, None]
        n_cut = float(((d2 <= pot.max_dist ** 2) & live).sum().item() / ns)
        del rows, d2, live
    else:
        t_nn = torch.from_numpy(sysm.numneigh).to(dev)
        t_neigh = torch.from_numpy(sysm.neigh).to(dev)
        t_off = torch.from_numpy(sysm.offsets).to(dev)
        list_stride = 0
        max_nn = int(sysm.numneigh[:nlocal].max())      # what a LAMMPS-KOKKOS list knows as d_neighbors.extent(1)
        # workload statistics for the roofline (listed / in-cutoff neighbors per atom)
        n_list = float(sysm.numneigh[:nlocal].mean())
        ii = np.repeat(np.arange(nlocal), sysm.numneigh[:nlocal])
        d2 = ((sysm.x[sysm.neigh] - sysm.x[ii]) ** 2).sum(axis=1)
        n_cut = float((d2 <= pot.max_dist ** 2).sum() / nlocal)
        del ii, d2
    flops_atom = algorithmic_flops_per_atom(pot, n_list, n_cut)
    bytes_atom = algorithmic_bytes_per_atom(n_list)
    t_f = torch.zeros((nall, 3), dtype=torch.float64, device=dev)
    t_ev = torch.zeros(8, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    t_grades = torch.zeros(nall, dtype=torch.float64, device=dev) if args.grade_every else None
    step_no = [0]

    lst = {"nn": t_nn, "neigh": t_neigh, "mx": max_nn}
    overlap = None
    if world > 1 and args.halo == "direct" and not args.no_overlap:
        overlap = decomp.OverlappedStep(halo, sysm.x[:nlocal], halo.sublo, halo.subhi, halo.rghost, dev)
        if not overlap.enabled:
            overlap = None

    def compute_part(il, evbuf, grade_step=False):
        mtp.compute_device(t_x, t_type, il, lst["nn"], lst["neigh"], t_off, t_f, evbuf, eflag=1, vflag=1,
                           variant=variant, stream=stream, max_numneigh=lst["mx"], grade=grade_step,
                           grades=t_grades if grade_step else None,
                           stride_i=int(lst["neigh"].shape[1]) if devlist_mode else 0, stride_jj=1)

    def compute_phased(il, counts, waits, dones, evbuf):
        mtp.compute_device_phased(counts, waits, dones, t_x, t_type, il, lst["nn"], lst["neigh"], t_off, t_f, evbuf, eflag=1,
                                  vflag=1, variant=variant, stream=stream, max_numneigh=lst["mx"],
                                  stride_i=int(lst["neigh"].shape[1]) if devlist_mode else 0, stride_jj=1)

    def step_device(rebuild=False):
        # what LAMMPS does around Pair::compute every step: forward comm of x, zero f, compute, reverse comm of f,
        # and the energy/virial all-reduce (rebuild: a re-neighboring step, list built on the device after the halo).
        # N > 1: the halo exchange runs behind the interior atoms (decomp.OverlappedStep)
        grade_step = bool(args.grade_every) and step_no[0] % args.grade_every == 0
        step_no[0] += 1
        if overlap is not None and not rebuild and not grade_step:
            overlap.run(t_x, t_f, t_ev, compute_part, t_ilist, compute_phased=compute_phased)
            return
        halo.forward(t_x)
        if rebuild:
            nn_new, tab, mx = mtp.neigh_build(t_x, nlocal, sysm.rlist, stream=stream)
            lst["nn"][:nlocal] = nn_new
            lst["neigh"], lst["mx"] = tab, mx
        t_f.zero_()
        compute_part(t_ilist, t_ev, grade_step)
        halo.reverse(t_f)
        halo.allreduce_ev(t_ev, grade_step)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    peaks = api.fp64_peaks(local_rank)       # measured DFMA / DMMA roofs (also warms the clocks)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    launches0 = api.kernel_launch_count() + halo.launches
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        flush.fill_(s & 0xff)                  # L2 flush between timed iterations (not timed)
        ev0[s].record()
        step_device()
        ev1[s].record()
    barrier()
    launches = api.kernel_launch_count() + halo.launches - launches0
    # per-kernel durations for the roofline: the same steps once more with the kernels serialised on one stream
    # (lanes = 1), CUDA events recorded by the library around every launch; not part of `value`
    prof_steps = min(args.steps, 5)
    mtp.set_lanes(1)
    step_device()
    mtp.profile_enable(True)
    for s in range(prof_steps):
        flush.fill_(s & 0xff)
        step_device()
    prof = mtp.profile_read()
    mtp.set_lanes(args.lanes)
    step_device()
    mtp.profile_read()
    for s in range(prof_steps):
        flush.fill_(s & 0xff)
        step_device()
    prof_lanes = mtp.profile_read()      # same events with the lanes of `value`: spans of different lanes overlap
    mtp.profile_enable(False)
    step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    mtp.synchronize()
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    value = world * nlocal / (ms_per_step * 1e-3) / 1e6
    energy = float(t_ev[0].item())

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, copies inside the timed region)
    e2e = None
    devlist = None
    if devlist_mode:
        # large systems: the list never exists on the host.  Per rank and step: H2D of the owned positions (and types on
        # re-neighboring steps) from pinned memory, halo forward, list rebuilt on the device (mtp_neigh_build), kernels,
        # halo reverse, EV all-reduce, D2H of the owned forces + EV record
        hx_t, _ = pinned_like(sysm.x[:nlocal])
        ht_t, _ = pinned_like(sysm.type)
        hf_t = torch.empty((nlocal, 3), dtype=torch.float64, pin_memory=True)
        hev_t = torch.empty(8, dtype=torch.float64, pin_memory=True)

        def step_e2e(k=0, list_every=1):
            t_x[:nlocal].copy_(hx_t, non_blocking=True)
            relist = k % list_every == 0
            if relist:
                t_type.copy_(ht_t, non_blocking=True)
            step_device(rebuild=relist)
            hf_t.copy_(t_f[:nlocal], non_blocking=True)
            hev_t.copy_(t_ev, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            step_e2e()
        barrier()
        e2e_n = {}
        for every in (1, LIST_EVERY):
            t0 = time.perf_counter()
            for k in range(args.steps):
                step_e2e(k, every)
            barrier()
            ms = 1e3 * (time.perf_counter() - t0) / args.steps
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            e2e_n[every] = ms
        e2e_ms, e2e10_ms = e2e_n[1], e2e_n[LIST_EVERY]
        list_bytes = 4 * nall
        h2d = 24 * nlocal + list_bytes
        d2h = 24 * nlocal + 64
        e2e_energy = float(hev_t[0])
        note = ("per rank and step, re-neighboring on EVERY step (worst case): H2D owned x, type from pinned memory, device halo "
                "forward, full neighbor list rebuilt on the device (mtp_neigh_build), kernels, halo reverse, EV all-reduce, D2H "
                "owned f + EV record; wall clock, max over ranks; bytes are per rank")
    elif world == 1:
        # N = 1: the reference-facing C-ABI call with HOST buffers (what a host-resident LAMMPS hands the pair style)
        keep, hx = pinned_like(sysm.x)
        k2, htype = pinned_like(sysm.type)
        k3, hnn = pinned_like(sysm.numneigh)
        k4, hneigh = pinned_like(sysm.neigh)
        k5, hoff = pinned_like(sysm.offsets)
        k6, hil = pinned_like(sysm.ilist)
        res = api.HostResult(nall, mtp.info.coeff_count)
        k7, res.f = pinned_like(res.f)
        k8, res.ev = pinned_like(res.ev)
        # f_overwrite: LAMMPS's force_clear zeroes f before Pair::compute, and a lone pair style is the first contributor
        # (PairMTPB200::compute sets the flag when force->pair is this style): the result is stored, f is not uploaded
        for _ in range(2):
            mtp.compute_host(hx, htype, hil, hnn, hneigh, hoff, eflag=1, vflag=1, variant=variant, out=res, f_overwrite=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            mtp.compute_host(hx, htype, hil, hnn, hneigh, hoff, eflag=1, vflag=1, variant=variant, out=res,
                             list_changed=True, f_overwrite=True)
        e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        # LAMMPS re-neighbors every ~10 steps at most; between rebuilds the list and the types stay resident on the device
        t0 = time.perf_counter()
        for k in range(args.steps):
            mtp.compute_host(hx, htype, hil, hnn, hneigh, hoff, eflag=1, vflag=1, variant=variant, out=res,
                             list_changed=(k % LIST_EVERY == 0), f_overwrite=True)
        e2e10_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        # informational: the list never crosses PCIe -- built on the device (mtp_neigh_build) on
        # every 10th step from the positions that were just uploaded; per step H2D x/type, kernels, D2H f + EV record
        try:
            hf_p = torch.empty((nall, 3), dtype=torch.float64, pin_memory=True)
            hev_p = torch.empty(8, dtype=torch.float64, pin_memory=True)
            hx_p, ht_p = keep, k2      # pinned torch tensors behind hx / htype
            t_nn2 = torch.zeros(nall, dtype=torch.int32, device=dev)
            state = {}

            def rebuild():
                nn_d, tab_d, mx_d = mtp.neigh_build(t_x, nlocal, sysm.rlist, stream=stream)
                t_nn2[:nlocal] = nn_d
                state.update(tab=tab_d, mx=mx_d)

            def step_devlist(k):
                t_x.copy_(hx_p, non_blocking=True)
                t_type.copy_(ht_p, non_blocking=True)
                if k % 10 == 0:
                    rebuild()
                t_f.zero_()
                tab_d = state["tab"]
                mtp.compute_device(t_x, t_type, t_ilist, t_nn2, tab_d, None, t_f, t_ev, stride_i=tab_d.shape[1], stride_jj=1,
                                   eflag=1, vflag=1, variant=variant, stream=stream, max_numneigh=state["mx"])
                hf_p.copy_(t_f, non_blocking=True)
                hev_p.copy_(t_ev, non_blocking=True)
                torch.cuda.synchronize()

            for k in range(2):
                step_devlist(0)
            tb0 = time.perf_counter()
            for _ in range(3):
                rebuild()
            torch.cuda.synchronize()
            build_ms = 1e3 * (time.perf_counter() - tb0) / 3
            t0 = time.perf_counter()
            for k in range(args.steps):
                step_devlist(k)
            dl_ms = 1e3 * (time.perf_counter() - t0) / args.steps
            same_list = bool(int(state["mx"]) == max_nn and
                             np.array_equal(t_nn2[:nlocal].cpu().numpy(), sysm.numneigh[:nlocal]))
            devlist = {"value": nlocal / (dl_ms * 1e-3) / 1e6, "ms_per_step": dl_ms, "neigh_build_ms": build_ms,
                       "h2d_bytes_per_step": int(28 * nall), "d2h_bytes_per_step": int(24 * nall + 64),
                       "neighbor_counts_match_host_list": same_list,
                       "energy_matches_device_path": bool(abs(float(hev_p[0]) - energy) <= 1e-9 * abs(energy)),
                       "note": "informational: H2D x/type every step, full neighbor list built on the device every 10th "
                               "step (mtp_neigh_build), kernels, D2H f + EV record; wall clock"}
        except Exception as exc:      # the headline e2e above does not depend on this variant
            devlist = {"error": repr(exc)}
        nid = nlocal
        list_bytes = 4 * sysm.neigh.size + 8 * nid + 4 * nid + 4 * nlocal + 4 * nall      # list + offsets + numneigh + ilist + type
        h2d = 24 * nall + list_bytes
        d2h = 24 * nall + 64
        e2e_energy = float(res.ev[0])
        note = ("mtp_compute_host per step with the neighbor list re-sent on EVERY step (worst case: LAMMPS re-neighbors far less "
                "often): H2D x, type, full neighbor list; kernels; D2H f + energy/virial record; f_overwrite (f is zero on entry, "
                "as after LAMMPS's force_clear); host buffers pinned; wall clock around the blocking call")
    else:
        # N > 1: same metric through the public Python API: per step H2D of the owned positions, types and the
        # neighbor list from pinned host memory, device halo exchange, kernels, D2H of the owned forces + EV record
        hx_t, _ = pinned_like(sysm.x[:nlocal])
        ht_t, _ = pinned_like(sysm.type)
        hnn_t, _ = pinned_like(sysm.numneigh)
        hne_t, _ = pinned_like(sysm.neigh)
        hof_t, _ = pinned_like(sysm.offsets)
        hf_t = torch.empty((nlocal, 3), dtype=torch.float64, pin_memory=True)
        hev_t = torch.empty(8, dtype=torch.float64, pin_memory=True)

        def step_e2e(k=0, list_every=1):
            t_x[:nlocal].copy_(hx_t, non_blocking=True)
            if k % list_every == 0:
                t_type.copy_(ht_t, non_blocking=True)
            if k % list_every == 0:      # re-neighboring step: the list crosses PCIe again
                t_nn.copy_(hnn_t, non_blocking=True)
                t_neigh.copy_(hne_t, non_blocking=True)
                t_off.copy_(hof_t, non_blocking=True)
            step_device()
            hf_t.copy_(t_f[:nlocal], non_blocking=True)
            hev_t.copy_(t_ev, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            step_e2e()
        barrier()
        e2e_n = {}
        for every in (1, LIST_EVERY):
            t0 = time.perf_counter()
            for k in range(args.steps):
                step_e2e(k, every)
            barrier()
            ms = 1e3 * (time.perf_counter() - t0) / args.steps
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_n[every] = float(t.item())
        e2e_ms, e2e10_ms = e2e_n[1], e2e_n[LIST_EVERY]
        list_bytes = 4 * nall + 4 * nall + 4 * sysm.neigh.size + 8 * (nall + 1)
        h2d = 24 * nlocal + list_bytes
        d2h = 24 * nlocal + 64
        e2e_energy = float(hev_t[0])
        note = ("per rank and step, neighbor list re-sent on EVERY step (worst case): H2D owned x, type, numneigh/offsets/neighbor "
                "list from pinned memory, device halo forward (NCCL send/recv), kernels, halo reverse, EV all-reduce, D2H owned "
                "f + EV record; wall clock, max over ranks; bytes are per rank")
    e2e = {"value": world * nlocal / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "note": note,
           "energy_matches_device_path": bool(abs(e2e_energy - energy) <= 1e-9 * abs(energy))}
    e2e["list_resident"] = {"value": world * nlocal / (e2e10_ms * 1e-3) / 1e6, "ms_per_step": e2e10_ms,
                            "h2d_bytes_per_step": int(h2d - list_bytes + list_bytes // LIST_EVERY), "d2h_bytes_per_step": int(d2h),
                            "note": "same step with the list (and types) re-sent on every %dth step only -- LAMMPS's "
                                    "re-neighboring cadence is at most that in a solid with the default 2 A skin" % LIST_EVERY}
    if devlist is not None:
        e2e["device_built_list"] = devlist

    # ---- informational: the reference's example deck as a device-resident MD loop (README.md:148-149: velocity
    # create + fix nve around the pair style): integrate, ghost images, list rebuild, pair style, reverse halo, all on
    # the device.  The random-init potential is rescaled by one factor to an RMS force of 0.05 eV/A.
    md = None
    if world == 1 and args.md_steps > 0 and not devlist_mode:
        try:
            from mtp_b200.md import NVE, scale_to_rms_force
            rms = float((t_f[:nlocal] ** 2).sum(dim=1).mean().sqrt().item())
            pot_md = scale_to_rms_force(pot, rms, 0.05)
            md_path = os.path.join(tmp, f"config{args.config}_md.almtp")
            almtp.write_almtp(md_path, pot_md)
            mtp_md = MTPB200(md_path, selection_state=False, device=local_rank)
            mtp_md.set_chunksize(args.chunksize)
            mtp_md.set_lanes(args.lanes)
            nve = NVE(mtp_md, sysm, halo, masses=cfg["masses"], dt=0.001, temperature=300.0, seed=12345, variant=variant)
            e_start = nve.potential_energy() + nve.kinetic_energy()
            nve.run(3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nve.run(args.md_steps)
            torch.cuda.synchronize()
            md_ms = 1e3 * (time.perf_counter() - t0) / args.md_steps
            e_end = nve.potential_energy() + nve.kinetic_energy()
            md = {"value": nlocal / (md_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": md_ms, "steps": args.md_steps,
                  "timestep_fs": 1.0, "temperature_K": nve.temperature(), "list_rebuilds": nve.rebuilds,
                  "total_energy_drift_eV_per_atom": abs(e_end - e_start) / nlocal,
                  "kinetic_energy_eV_per_atom": nve.kinetic_energy() / nlocal,
                  "note": "informational: NVE velocity Verlet, every stage of the step on the device (mtp_nve_*_integrate, ghost "
                          "images, mtp_neigh_build when an atom moved more than half the skin, mtp_compute, reverse halo); "
                          "wall clock, one device sync per 10 steps"}
            mtp_md.close()
        except Exception as exc:
            md = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    dfma, dmma = peaks
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    split = algorithmic_flops_split(pot, n_list, n_cut)
    try:
        cg_info = api.codegen_source(pot_path, args.variant == "small")[1]
        program_terms = cg_info["terms"]
    except Exception:
        cg_info, program_terms = None, 3 * pot.T
    executed = executed_flops_split(pot, n_list, n_cut, program_terms)

    def kernel_table(p):
        out = {}
        for cls, (ms, spans) in p.items():
            if spans == 0:
                continue
            k = {"ms_per_step": ms / prof_steps, "launches_per_step": spans / prof_steps}
            if cls in split:
                k["algorithmic_flops_per_atom"] = split[cls]
                k["achieved_tflops"] = split[cls] * nlocal / (ms / prof_steps * 1e-3) / 1e12
                k["frac_of_fp64_peak"] = k["achieved_tflops"] / dfma
                k["executed_flops_per_atom"] = executed[cls]
                k["executed_tflops"] = executed[cls] * nlocal / (ms / prof_steps * 1e-3) / 1e12
                k["executed_frac_of_fp64_peak"] = k["executed_tflops"] / dfma
            out[cls] = k
        return out

    kernels = kernel_table(prof)
    kernels_lanes = kernel_table(prof_lanes)
    dom = max((c for c in kernels if c in split), key=lambda c: kernels[c]["ms_per_step"], default=None)
    pipe_ms = sum(k["ms_per_step"] for k in kernels.values())
    achieved_tf = flops_atom * nlocal / (pipe_ms * 1e-3) / 1e12 if pipe_ms else 0.0
    executed_atom = sum(executed.values())
    achieved_gbs = bytes_atom * nlocal / (ms_per_step * 1e-3) / 1e9
    if "program" in kernels and cg_info:
        # the generated program kernel is straight-line code executed once per 32-atom chunk: it streams through the
        # instruction cache and is bound by instruction fetch (ncu: ~0.5 instructions/clk/SM, profiles/r2_*), not by
        # the FP64 or the shared-memory pipe; reported for what it is
        kernels["program"]["generated"] = {k: cg_info[k] for k in ("atoms_per_cta", "warps", "ctas_per_sm", "rows", "stages",
                                                                   "smem_bytes", "terms", "loads", "stores")}
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture (config 2 only: the
    # capture is of that workload, bench defaults); null for any other workload
    traffic, traffic_note = None, None
    try:
        if args.config == 2 and not args.cells and dom:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tj["kernels"][dom]["bytes_per_launch"]
            traffic_note = tj["note"]
    except Exception:
        traffic = None
    roofline = {"bound": "fp64", "unit": "TFLOP/s", "kernel": dom,
                "achieved": kernels[dom]["achieved_tflops"] if dom else None, "peak": dfma,
                "frac": kernels[dom]["frac_of_fp64_peak"] if dom else None,
                "achieved_executed": kernels[dom]["executed_tflops"] if dom else None,
                "frac_executed": kernels[dom]["executed_frac_of_fp64_peak"] if dom else None,
                "frac_note": "dominant kernel (largest share of the step) on the FP64 roof: `frac` counts the flops as written in "
                             "pair_mtp.cpp (SURVEY.md 8d, includes forming the Jacobian, which this design never does -- it can "
                             "exceed 1), `frac_executed` counts the FP64 operations the kernel actually executes",
                "traffic": traffic, "traffic_note": traffic_note,
                "duration_ms": kernels[dom]["ms_per_step"] / kernels[dom]["launches_per_step"] if dom else None,
                "duration_note": "average launch duration of the dominant kernel: CUDA events recorded by the library on its "
                                 "launch stream around every launch (mtp_profile_read), kernels serialised (lanes = 1) in a pass "
                                 "right after the timed region; `kernels_lanes_on` holds the same events with the lanes of `value` "
                                 "(spans of different lanes overlap there, so they sum to more than the step)",
                "peak_source": "FP64 DFMA peak measured in this run by mtp_fp64_peak (DMMA m8n8k4: %.2f TFLOP/s); "
                               "MEASURED_PEAKS.json has no FP64 entry" % dmma,
                "kernels": kernels, "kernels_lanes_on": kernels_lanes,
                "pipeline": {"ms_per_step_kernels": pipe_ms, "algorithmic_flops_per_atom": flops_atom,
                             "achieved": achieved_tf, "frac": achieved_tf / dfma,
                             "executed_flops_per_atom": executed_atom,
                             "frac_executed": executed_atom * nlocal / (pipe_ms * 1e-3) / 1e12 / dfma if pipe_ms else None,
                             "frac_of_value": flops_atom * nlocal / (ms_per_step * 1e-3) / 1e12 / dfma,
                             "note": "all kernels of one force evaluation, serialised (`frac`, `frac_executed`) and as timed for "
                                     "`value` with the lanes on (`frac_of_value`); flops counted as written in pair_mtp.cpp "
                                     "(SURVEY.md 8d) unless marked executed"},
                "algorithmic_bytes_per_atom": bytes_atom, "neighbors_listed": n_list, "neighbors_in_cutoff": n_cut,
                "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                        "peak_source": hbm_src}}

    cpu = None
    if not args.no_cpu_baseline:
        r = run_reference(argparse.Namespace(**{**vars(args), "steps": 8, "warmup": 1}), cfg, pot_path, pot)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
               "cpu_model": cpu_model()}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "atoms_per_gpu": nlocal, "ghosts_per_gpu": nall - nlocal,
                       "parallelism": "brick grid %dx%dx%d, one rank per GPU, %s NCCL send/recv halo (%d B/rank/step)%s" % (
                           *brick_grid(world), args.halo, halo.bytes_per_step,
                           ", exchange overlapped with the interior atoms (%d + %d interior, %d boundary centres)" % (
                               overlap.counts[0], overlap.counts[2], overlap.counts[1]) if overlap is not None else ""),
                       "l2": "256 MiB write between timed iterations (L2 flush), per-step CUDA events summed",
                       "variant": args.variant, "chunksize": args.chunksize, "lanes": args.lanes, "flags": "eflag=1 vflag=1",
                       "neighbor_list": "built on the device (mtp_neigh_build)" if devlist_mode else "built on the host, resident on the device",
                       "grade_every": args.grade_every},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "energy": energy}
    if md is not None:
        line["md"] = md
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
