"""Brick spatial decomposition + ghost-atom halo exchange, one rank per GPU.

Stands in for what upstream LAMMPS's ``Comm`` does around the pair style (pair_mtp.cpp:81-88,113,315
rely on it: only ``inum`` owned centres are evaluated, ghosts are reached through the neighbor list and
their forces are returned with ``newton on``):

* setup: LAMMPS's six-swap scheme -- per dimension, atoms within ``r_cut + skin`` of the low / high face
  (including ghosts received in earlier dimensions, so edges and corners come for free) are sent to the
  minus / plus neighbor brick, shifted by the global box length when they cross the periodic boundary;
* every step: ``forward`` packs ghost-source positions on the device (``mtp_halo_pack_x``), exchanges
  them with grouped NCCL send/recv over NVLink and lands them directly in the ghost rows of ``x``;
  ``reverse`` sends the ghost rows of ``f`` back and scatter-adds them (``mtp_halo_unpack_add_f``);
  energies/virials need one 7-double all-reduce (``allreduce_ev``).

The transport is ``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests, where pack/unpack run as
torch index ops -- host-logic coverage only, the product path is the CUDA one).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import harness


def brick_grid(n: int):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


def rank_coords(rank: int, grid):
    return (rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1]))


def coords_rank(c, grid):
    return (c[0] % grid[0]) + grid[0] * ((c[1] % grid[1]) + grid[1] * (c[2] % grid[2]))


@dataclass
class Swap:
    dim: int
    peer_send: int            # rank the packed rows go to
    peer_recv: int            # rank the ghost rows come from
    sendlist: torch.Tensor    # int32 [n_send] rows of x to pack
    shift: np.ndarray         # [3] periodic shift added while packing
    recv_first: int           # first ghost row filled by this swap
    recv_n: int
    sendbuf: torch.Tensor | None = None     # [n_send, 3] f64
    recvbuf: torch.Tensor | None = None     # [recv_n, 3] f64 (reverse direction)


class Halo:
    """Per-step forward (x) / reverse (f) ghost exchange for one rank."""

    def __init__(self, swaps, rank, world, device, lib=None):
        self.swaps = swaps
        self.rank, self.world, self.device = rank, world, device
        self.lib = lib            # ctypes handle of libmtp_b200.so when on CUDA
        self.cuda = device.type == "cuda"
        for s in swaps:
            s.sendbuf = torch.empty((len(s.sendlist), 3), dtype=torch.float64, device=device)
            s.recvbuf = torch.empty((len(s.sendlist), 3), dtype=torch.float64, device=device)
        self.bytes_per_step = sum(2 * 24 * len(s.sendlist) for s in swaps if s.peer_send != rank)
        self.launches = 0

    # -- pack / unpack: CUDA kernels of the C ABI on the GPU, index ops on the CPU (tests only)
    def _pack(self, x, s: Swap, out):
        n = len(s.sendlist)
        if n == 0:
            return
        if self.cuda:
            import ctypes as C
            sh = (C.c_double * 3)(*s.shift)
            rc = self.lib.mtp_halo_pack_x(x.data_ptr(), s.sendlist.data_ptr(), n, sh, out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream)
            assert rc == 0
            self.launches += 1
        else:
            out.copy_(x.index_select(0, s.sendlist.long()) + torch.from_numpy(s.shift))

    def _unpack_add(self, f, s: Swap, buf):
        n = len(s.sendlist)
        if n == 0:
            return
        if self.cuda:
            rc = self.lib.mtp_halo_unpack_add_f(f.data_ptr(), s.sendlist.data_ptr(), n, buf.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream)
            assert rc == 0
            self.launches += 1
        else:
            f.index_add_(0, s.sendlist.long(), buf.clone())

    def _by_dim(self):
        for d in range(3):
            grp = [s for s in self.swaps if s.dim == d]
            if grp:
                yield grp

    def forward(self, x: torch.Tensor):
        """Ghost rows of x <- owners' current positions (LAMMPS Comm::forward_comm)."""
        for grp in self._by_dim():
            ops, reqs = [], []
            for s in grp:
                ghost = x[s.recv_first: s.recv_first + s.recv_n]
                if s.peer_send == self.rank:
                    self._pack(x, s, ghost)           # periodic self-image: pack straight into the ghost rows
                else:
                    self._pack(x, s, s.sendbuf)
            for s in grp:
                if s.peer_send != self.rank and len(s.sendlist):
                    ops.append(dist.P2POp(dist.isend, s.sendbuf, s.peer_send))
            for s in grp:
                if s.peer_recv != self.rank and s.recv_n:
                    ops.append(dist.P2POp(dist.irecv, x[s.recv_first: s.recv_first + s.recv_n], s.peer_recv))
            if ops:
                reqs = dist.batch_isend_irecv(ops)
                for r in reqs:
                    r.wait()

    def reverse(self, f: torch.Tensor):
        """Owners' f += ghost rows of f (LAMMPS Comm::reverse_comm, newton on), swaps in reverse order."""
        for grp in reversed(list(self._by_dim())):
            ops = []
            for s in grp:
                if s.peer_recv != self.rank and s.recv_n:
                    ops.append(dist.P2POp(dist.isend, f[s.recv_first: s.recv_first + s.recv_n], s.peer_recv))
            for s in grp:
                if s.peer_send != self.rank and len(s.sendlist):
                    ops.append(dist.P2POp(dist.irecv, s.recvbuf, s.peer_send))
            if ops:
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
            for s in grp:
                if s.peer_send == self.rank:
                    self._unpack_add(f, s, f[s.recv_first: s.recv_first + s.recv_n])
                else:
                    self._unpack_add(f, s, s.recvbuf)

    def allreduce_ev(self, ev: torch.Tensor, grade: bool = True):
        """E + 6 virial components: SUM over ranks; ev[7] (max grade, grade steps only): MAX
        (pair_mtp_extrapolation.cpp:379).  Neighbourhood mode only: in configuration mode the candidate VECTOR is summed
        over ranks and the grade re-evaluated (pair_mtp_extrapolation.cpp:366-376), see allreduce_cfg_grade."""
        if self.world > 1:
            dist.all_reduce(ev[:7], op=dist.ReduceOp.SUM)
            if grade:
                dist.all_reduce(ev[7:8], op=dist.ReduceOp.MAX)


def _exchange_arrays(send_to, recv_from, arrays, device, rank):
    """Setup-time exchange of numpy arrays with two peers (one op group): returns what recv_from sent."""
    metas = [(a.shape, str(a.dtype)) for a in arrays]
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, (rank, send_to, metas))
    peer_metas = None
    for (r, dst, m) in gathered:
        if r == recv_from and dst == rank:
            peer_metas = m
    assert peer_metas is not None
    sends = [torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in arrays]
    recvs = [torch.empty(shape, dtype=getattr(torch, {"float64": "float64", "int32": "int32"}[dt]), device=device)
             for (shape, dt) in peer_metas]
    ops = [dist.P2POp(dist.isend, t, send_to) for t in sends if t.numel()]
    ops += [dist.P2POp(dist.irecv, t, recv_from) for t in recvs if t.numel()]
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    if device.type == "cuda":
        torch.cuda.synchronize()
    return [t.cpu().numpy() for t in recvs]


def build_rank_system(x_own, types_own, sublo, subhi, grid, rank, gbox, cutoff, skin, device, lib=None, with_list=True):
    """Ghost shell + send lists for this rank's brick [sublo, subhi) of the periodic global box ``gbox``."""
    world = grid[0] * grid[1] * grid[2]
    me = rank_coords(rank, grid)
    rghost = cutoff + skin
    x = np.ascontiguousarray(x_own, dtype=np.float64)
    t = np.ascontiguousarray(types_own, dtype=np.int32)
    nlocal = len(x)
    swaps = []
    for d in range(3):
        if rghost >= (subhi[d] - sublo[d]):
            raise ValueError("ghost cutoff exceeds the brick length; use a larger per-GPU box")
        lo = np.nonzero(x[:, d] < sublo[d] + rghost)[0].astype(np.int32)
        hi = np.nonzero(x[:, d] >= subhi[d] - rghost)[0].astype(np.int32)
        sh_lo, sh_hi = np.zeros(3), np.zeros(3)
        if me[d] == 0:
            sh_lo[d] = gbox[d]
        if me[d] == grid[d] - 1:
            sh_hi[d] = -gbox[d]
        cm, cp = list(me), list(me)
        cm[d] -= 1
        cp[d] += 1
        minus, plus = coords_rank(cm, grid), coords_rank(cp, grid)
        first = len(x)
        # this rank's lo-send lands on `minus`; what lands here first comes from `plus` (its lo-send), then
        # from `minus` (its hi-send)
        if grid[d] == 1:
            got = [(x[lo] + sh_lo, t[lo]), (x[hi] + sh_hi, t[hi])]
        else:
            a = _exchange_arrays(minus, plus, [x[lo] + sh_lo, t[lo]], device, rank)
            b = _exchange_arrays(plus, minus, [x[hi] + sh_hi, t[hi]], device, rank)
            got = [(a[0].reshape(-1, 3), a[1]), (b[0].reshape(-1, 3), b[1])]
        n_a, n_b = len(got[0][0]), len(got[1][0])
        swaps.append(Swap(d, minus, plus, torch.from_numpy(lo).to(device), sh_lo, first, n_a))
        swaps.append(Swap(d, plus, minus, torch.from_numpy(hi).to(device), sh_hi, first + n_a, n_b))
        x = np.concatenate([x, got[0][0], got[1][0]])
        t = np.concatenate([t, got[0][1], got[1][1]])
    x = np.ascontiguousarray(x)
    t = np.ascontiguousarray(t.astype(np.int32))
    # with_list=False: the caller builds the list on the device (mtp_neigh_build); host arrays stay empty
    numneigh, offsets, flat = harness.neighbor_list(x, nlocal, rghost) if with_list else (
        np.zeros(len(x), np.int32), np.zeros(len(x) + 1, np.int64), np.zeros(0, np.int32))
    owner = np.full(len(x), -1, dtype=np.int32)
    owner[:nlocal] = np.arange(nlocal, dtype=np.int32)
    sysm = harness.System(box=np.asarray(gbox, dtype=np.float64), nlocal=nlocal, x=x, type=t, owner=owner,
                          ilist=np.arange(nlocal, dtype=np.int32), numneigh=numneigh, offsets=offsets, neigh=flat,
                          rlist=rghost)
    return sysm, Halo(swaps, rank, world, device, lib)


DIRS = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1) if (dx, dy, dz) != (0, 0, 0)]


class DirectHalo:
    """Single-stage ghost exchange: every rank sends its owned boundary atoms straight to each of the 26
    neighbor directions (faces, edges, corners), so a step needs ONE device pack, ONE grouped NCCL send/recv
    each way and ONE scatter-add, instead of three dependent stages.  Ghost rows are laid out as
    [self-image segments | remote segments], both in ``DIRS`` order; two ranks that exchange several segments
    enumerate them in the same order, which is what NCCL's in-order matching between a pair needs."""

    def __init__(self, rank, world, device, lib, nlocal, segs):
        # segs: dicts with dir index, dest, src, sendlist (np int32), shift (3,), recv_n
        self.rank, self.world, self.device, self.lib = rank, world, device, lib
        self.cuda = device.type == "cuda"
        self.launches = 0
        self.nlocal = nlocal
        selfs = [s for s in segs if s["dest"] == rank]
        # remote segments grouped by peer (DIRS order inside a peer): one message per peer and direction of travel
        rem = sorted([s for s in segs if s["dest"] != rank], key=lambda s: (s["dest"], s["dir"]))
        rem_in = sorted([s for s in segs if s["src"] != rank], key=lambda s: (s["src"], s["dir"]))
        self.remote = rem

        def cat(lst, key, dtype):
            return np.concatenate([s[key] for s in lst] + [np.zeros(0, dtype)]).astype(dtype)

        def seg_ids(lst):
            return np.concatenate([np.full(len(s["sendlist"]), k, np.uint8) for k, s in enumerate(lst)]
                                  + [np.zeros(0, np.uint8)])

        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        self.self_list, self.self_seg = t(cat(selfs, "sendlist", np.int32)), t(seg_ids(selfs))
        self.self_shift = t(np.array([s["shift"] for s in selfs] + [[0, 0, 0]], dtype=np.float64))
        self.rem_list, self.rem_seg = t(cat(rem, "sendlist", np.int32)), t(seg_ids(rem))
        self.rem_shift = t(np.array([s["shift"] for s in rem] + [[0, 0, 0]], dtype=np.float64))
        self.nself = len(self.self_list)
        # per peer: [offset, count] of its block in the send buffer and in the ghost rows
        self.send_blocks, self.recv_blocks = {}, {}
        so = 0
        for s in rem:
            b = self.send_blocks.setdefault(s["dest"], [so, 0])
            b[1] += len(s["sendlist"])
            so += len(s["sendlist"])
        g = nlocal + self.nself
        for s in rem_in:
            b = self.recv_blocks.setdefault(s["src"], [g, 0])
            b[1] += s["recv_n"]
            g += s["recv_n"]
        self.nall = g
        self.ghost_order = [s["dir"] for s in selfs] + [s["dir"] for s in rem_in]
        self.sendbuf = torch.empty((so, 3), dtype=torch.float64, device=device)
        self.recvbuf = torch.empty((so, 3), dtype=torch.float64, device=device)
        self.bytes_per_step = 2 * 24 * so
        self._ops = {}

    def _pack(self, x, lst, seg, shifts, out):
        n = len(lst)
        if n == 0:
            return
        if self.cuda:
            rc = self.lib.mtp_halo_pack_x_multi(x.data_ptr(), lst.data_ptr(), seg.data_ptr(), shifts.data_ptr(), n,
                                                out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            assert rc == 0
            self.launches += 1
        else:
            out.copy_(x.index_select(0, lst.long()) + shifts.index_select(0, seg.long()))

    def _unpack_add(self, f, lst, buf):
        n = len(lst)
        if n == 0:
            return
        if self.cuda:
            rc = self.lib.mtp_halo_unpack_add_f(f.data_ptr(), lst.data_ptr(), n, buf.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream)
            assert rc == 0
            self.launches += 1
        else:
            f.index_add_(0, lst.long(), buf.clone())

    # The exchange is split in two halves so that a caller can put work between them: *_begin packs and posts the grouped
    # NCCL send/recv (which run on NCCL's own stream), *_end makes the current stream wait for them and unpacks.
    def forward_begin(self, x: torch.Tensor):
        n0 = self.nlocal
        self._pack(x, self.self_list, self.self_seg, self.self_shift, x[n0: n0 + self.nself])
        self._pack(x, self.rem_list, self.rem_seg, self.rem_shift, self.sendbuf)
        key = ("fwd", x.data_ptr())
        ops = self._ops.get(key)
        if ops is None:      # the P2P descriptors only depend on the buffers: built once per x tensor
            ops = [dist.P2POp(dist.isend, self.sendbuf[o: o + n], p) for p, (o, n) in sorted(self.send_blocks.items()) if n]
            ops += [dist.P2POp(dist.irecv, x[o: o + n], p) for p, (o, n) in sorted(self.recv_blocks.items()) if n]
            self._ops[key] = ops
        return dist.batch_isend_irecv(ops) if ops else []

    def forward_end(self, works):
        for r in works:
            r.wait()

    def forward(self, x: torch.Tensor):
        self.forward_end(self.forward_begin(x))

    def reverse_begin(self, f: torch.Tensor):
        key = ("rev", f.data_ptr())
        ops = self._ops.get(key)
        if ops is None:
            ops = [dist.P2POp(dist.isend, f[o: o + n], p) for p, (o, n) in sorted(self.recv_blocks.items()) if n]
            ops += [dist.P2POp(dist.irecv, self.recvbuf[o: o + n], p) for p, (o, n) in sorted(self.send_blocks.items()) if n]
            self._ops[key] = ops
        return dist.batch_isend_irecv(ops) if ops else []

    def reverse_end(self, f: torch.Tensor, works):
        for r in works:
            r.wait()
        n0 = self.nlocal
        self._unpack_add(f, self.self_list, f[n0: n0 + self.nself])
        self._unpack_add(f, self.rem_list, self.recvbuf)

    def reverse(self, f: torch.Tensor):
        self.reverse_end(f, self.reverse_begin(f))

    def allreduce_ev(self, ev: torch.Tensor, grade: bool = True):
        """E + 6 virial components: one SUM over ranks; the max grade ev[7] (pair_mtp_extrapolation.cpp:379) only on
        grade steps."""
        if self.world > 1:
            dist.all_reduce(ev[:7], op=dist.ReduceOp.SUM)
            if grade:
                dist.all_reduce(ev[7:8], op=dist.ReduceOp.MAX)


def allreduce_cfg_grade(candidate: torch.Tensor, inverse_active_set: torch.Tensor, natoms_total: int) -> torch.Tensor:
    """Configuration-mode grade over ranks (pair_mtp_extrapolation.cpp:366-376, ReduceCoeffDers + compile_grades of the
    KOKKOS styles): the per-rank candidate vectors [Q] are SUMMED (one all-reduce of Q doubles, in place), and the grade
    max_i |Ainv[i, :] . b| / natoms is evaluated from the sum on the device.  A MAX over per-rank grades -- what
    ``allreduce_ev`` does with ev[7] -- is only right in neighbourhood mode."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(candidate, op=dist.ReduceOp.SUM)
    return (inverse_active_set @ candidate).abs().max() / float(natoms_total)


def split_interior(x_own: np.ndarray, sublo, subhi, rghost: float):
    """Owned atoms whose whole neighbor list (radius rghost = cutoff + skin) lies inside the brick cannot have a ghost
    neighbor: their forces need no halo.  Returns (interior ids, boundary ids), both ascending int32."""
    lo, hi = np.asarray(sublo, dtype=np.float64), np.asarray(subhi, dtype=np.float64)
    inside = np.all((x_own > lo + rghost) & (x_own < hi - rghost), axis=1)
    ids = np.arange(len(x_own), dtype=np.int32)
    return ids[inside], ids[~inside]


class OverlappedStep:
    """One force evaluation of a rank with the halo exchange hidden behind the interior atoms:

        forward halo posted  ||  pair style on the first half of the interior atoms
        pair style on the boundary atoms (waits for the ghosts)
        reverse halo posted  ||  pair style on the second half of the interior atoms
        ghost forces added (atomic adds, concurrent with the interior forces), one 7-double all-reduce

    GPU: ONE ``mtp_compute_phased`` call -- the three runs of the list share the library's lanes, the boundary phase waits
    for a CUDA event recorded behind the NCCL receives, and the reverse exchange starts on a side stream from the event
    the library records when the boundary phase is complete.  CPU (gloo tests; ``compute`` = closure around the checker):
    three partial evaluations in the same order.  With one rank the plain sequence forward / compute / reverse is used."""

    def __init__(self, halo, x_own, sublo, subhi, rghost, device, min_part=16384):
        self.halo = halo
        self.device = device
        interior, boundary = split_interior(x_own, sublo, subhi, rghost)
        self.enabled = isinstance(halo, DirectHalo) and halo.world > 1 and len(interior) >= 2 * min_part and len(boundary) > 0
        half = len(interior) // 2
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        self.parts = [t(interior[:half]), t(boundary), t(interior[half:])] if self.enabled else []
        self.counts = (int(half), int(len(boundary)), int(len(interior) - half))
        self.ilist = torch.cat(self.parts) if self.enabled else None
        self.evs = [torch.zeros(8, dtype=torch.float64, device=device) for _ in range(3)]
        if self.enabled and device.type == "cuda":
            self.side = [torch.cuda.Stream(device=device) for _ in range(2)]
            self.ev_halo = torch.cuda.Event()
            self.ev_boundary = torch.cuda.Event()
            self.ev_reverse = torch.cuda.Event()

    def run(self, x, f, ev, compute, all_ilist, grade=False, compute_phased=None):
        h = self.halo
        if not self.enabled:
            h.forward(x)
            f.zero_()
            compute(all_ilist, ev)
            h.reverse(f)
            h.allreduce_ev(ev, grade)
            return
        if compute_phased is not None and self.device.type == "cuda":
            main = torch.cuda.current_stream()
            works = h.forward_begin(x)
            with torch.cuda.stream(self.side[0]):       # the event completes when the ghost positions have landed
                h.forward_end(works)
                self.ev_halo.record()
            f.zero_()
            compute_phased(self.ilist, self.counts, [None, self.ev_halo, None], [None, self.ev_boundary, None], ev)
            self.side[1].wait_event(self.ev_boundary)
            with torch.cuda.stream(self.side[1]):       # ghost forces travel back under the second interior phase
                h.reverse_end(f, h.reverse_begin(f))
                self.ev_reverse.record()
            main.wait_event(self.ev_reverse)
            h.allreduce_ev(ev, grade)
            return
        works = h.forward_begin(x)
        f.zero_()
        compute(self.parts[0], self.evs[0])
        h.forward_end(works)
        compute(self.parts[1], self.evs[1])
        works = h.reverse_begin(f)
        compute(self.parts[2], self.evs[2])
        h.reverse_end(f, works)
        torch.add(self.evs[0], self.evs[1], out=ev)
        ev.add_(self.evs[2])
        if grade:
            ev[7] = torch.maximum(torch.maximum(self.evs[0][7], self.evs[1][7]), self.evs[2][7])
        h.allreduce_ev(ev, grade)


def build_rank_system_direct(x_own, types_own, sublo, subhi, grid, rank, gbox, cutoff, skin, device, lib=None, with_list=True):
    """Same ghost shell as ``build_rank_system`` (possibly in a different row order), built for ``DirectHalo``."""
    world = grid[0] * grid[1] * grid[2]
    me = rank_coords(rank, grid)
    rghost = cutoff + skin
    x = np.ascontiguousarray(x_own, dtype=np.float64)
    t = np.ascontiguousarray(types_own, dtype=np.int32)
    nlocal = len(x)
    lo, hi = [], []
    for d in range(3):
        if 2 * rghost >= (subhi[d] - sublo[d]) and grid[d] <= 2:
            pass    # an atom may then be a ghost of the same neighbor through both faces; still correct
        if rghost >= (subhi[d] - sublo[d]):
            raise ValueError("ghost cutoff exceeds the brick length; use a larger per-GPU box")
        lo.append(x[:, d] < sublo[d] + rghost)
        hi.append(x[:, d] >= subhi[d] - rghost)
    segs = []
    for k, dv in enumerate(DIRS):
        sel = np.ones(nlocal, dtype=bool)
        shift = np.zeros(3)
        for d in range(3):
            if dv[d] == -1:
                sel &= lo[d]
                if me[d] == 0:
                    shift[d] = gbox[d]
            elif dv[d] == 1:
                sel &= hi[d]
                if me[d] == grid[d] - 1:
                    shift[d] = -gbox[d]
        dest = coords_rank([me[d] + dv[d] for d in range(3)], grid)
        src = coords_rank([me[d] - dv[d] for d in range(3)], grid)
        segs.append(dict(dir=k, dest=dest, src=src, sendlist=np.nonzero(sel)[0].astype(np.int32), shift=shift))
    # segment sizes of every rank -> what each of my directions receives
    counts = [len(s["sendlist"]) for s in segs]
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, counts)
    else:
        allc = [counts]
    for s in segs:
        s["recv_n"] = allc[s["src"]][s["dir"]]
    # setup exchange of ghost positions and types: one group, same matching order as the per-step exchange
    ghosts_x, ghosts_t = {}, {}
    sends, recvs, ops = [], [], []
    for s in segs:
        px, pt = x[s["sendlist"]] + s["shift"], t[s["sendlist"]]
        if s["dest"] == rank:
            ghosts_x[s["dir"]], ghosts_t[s["dir"]] = px, pt
        else:
            tx = torch.from_numpy(np.ascontiguousarray(px)).to(device)
            tt = torch.from_numpy(np.ascontiguousarray(pt)).to(device)
            sends.append((tx, tt))
            if tx.numel():
                ops += [dist.P2POp(dist.isend, tx, s["dest"]), dist.P2POp(dist.isend, tt, s["dest"])]
    for s in segs:
        if s["dest"] != rank:
            rx = torch.empty((s["recv_n"], 3), dtype=torch.float64, device=device)
            rt = torch.empty(s["recv_n"], dtype=torch.int32, device=device)
            recvs.append((s["dir"], rx, rt))
            if rx.numel():
                ops += [dist.P2POp(dist.irecv, rx, s["src"]), dist.P2POp(dist.irecv, rt, s["src"])]
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        if device.type == "cuda":
            torch.cuda.synchronize()
    for k, rx, rt in recvs:
        ghosts_x[k], ghosts_t[k] = rx.cpu().numpy(), rt.cpu().numpy()
    halo = DirectHalo(rank, world, device, lib, nlocal, segs)
    halo.sublo, halo.subhi, halo.rghost = np.asarray(sublo, dtype=np.float64), np.asarray(subhi, dtype=np.float64), rghost
    order = halo.ghost_order
    x = np.ascontiguousarray(np.concatenate([x] + [ghosts_x[k].reshape(-1, 3) for k in order]))
    t = np.ascontiguousarray(np.concatenate([t] + [ghosts_t[k] for k in order]).astype(np.int32))
    assert len(x) == halo.nall
    # with_list=False: the caller builds the list on the device (mtp_neigh_build); host arrays stay empty
    numneigh, offsets, flat = harness.neighbor_list(x, nlocal, rghost) if with_list else (
        np.zeros(len(x), np.int32), np.zeros(len(x) + 1, np.int64), np.zeros(0, np.int32))
    owner = np.full(len(x), -1, dtype=np.int32)
    owner[:nlocal] = np.arange(nlocal, dtype=np.int32)
    sysm = harness.System(box=np.asarray(gbox, dtype=np.float64), nlocal=nlocal, x=x, type=t, owner=owner,
                          ilist=np.arange(nlocal, dtype=np.int32), numneigh=numneigh, offsets=offsets, neigh=flat,
                          rlist=rghost)
    return sysm, halo


def make_rank_system(config, cells, grid, rank, device, lib=None, cutoff=5.0, skin=2.0, jitter=0.05, direct=False,
                     with_list=True):
    """BASELINE.json weak-scaling layout: the per-GPU box of ``config`` replicated on the brick grid."""
    cfg = harness.CONFIGS[config]
    x, box = harness.lattice(cfg["kind"], cfg["a"], cells or cfg["cells"], jitter=jitter, seed=2024)
    types = harness.random_types(x.shape[0], cfg["fractions"], cfg["type_seed"])
    me = np.array(rank_coords(rank, grid), dtype=np.float64)
    sublo = me * box
    build = build_rank_system_direct if direct else build_rank_system
    return build(x + sublo, types, sublo, sublo + box, grid, rank, box * np.array(grid), cutoff, skin, device, lib,
                 with_list=with_list)
