"""MTP level -> index tables (alpha_index_basic / alpha_index_times / alpha_moment_mapping).

No trained ``.almtp`` ships with the reference and the reference parser rejects
untrained templates (pair_mtp.cpp:434-439, :556-569), so the potentials used by
the tests and by ``bench.py`` are generated here.  The tables follow the MLIP
convention the reference consumes (pair_mtp.cpp:154-201):

* a basic moment ``k = (mu, ax, ay, az)`` is the (ax,ay,az) component of the rank
  ``nu = ax+ay+az`` moment tensor ``M_{mu,nu}``, level ``2 + 4 mu + nu``;
* ``alpha_index_times`` is a program of ``m[a3] += mult * m[a0] * m[a1]`` edges,
  topologically ordered and sorted by target, at most 3 dependency waves deep
  (pair_mtps_kokkos.cpp:179-200);
* a basis function is a complete contraction (loop-free multigraph) of at most
  four moment tensors with total level <= L  (this rule
  reproduces the known scalar counts 1,2,5,9,16,29,52,92,163,288,500,864 for
  L = 2..24).

Contractions are lowered to component products with multinomial multiplicities:
contracting ``n`` symmetric indices, ``sum_{a1..an} A_{a1..an} B_{a1..an} =
sum_{|g|=n} n!/(gx! gy! gz!) A_g B_g``.  Intermediates are memoised by a
canonical form so that shared sub-contractions (``M00^2``, ``M01.M01`` ...) are
built once, and only the components some basis function actually needs are
emitted.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from functools import lru_cache
from math import factorial


def moment_level(mu: int, nu: int) -> int:
    return 2 + 4 * mu + nu


@lru_cache(maxsize=None)
def multi_indices(n: int):
    """All (ax,ay,az) with ax+ay+az == n, ordered (n,0,0),(n-1,1,0),(n-1,0,1),... as in MLIP files."""
    out = []
    for ax in range(n, -1, -1):
        for ay in range(n - ax, -1, -1):
            out.append((ax, ay, n - ax - ay))
    return tuple(out)


def multinomial(g) -> int:
    return factorial(sum(g)) // (factorial(g[0]) * factorial(g[1]) * factorial(g[2]))


# --------------------------------------------------------------------------- graphs

def _row_sum_matrices(nus):
    """Symmetric non-negative integer matrices, zero diagonal, row sums == nus (upper triangles)."""
    n = len(nus)
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    res = []

    def rec(p, rem, cur):
        if p == len(pairs):
            if all(r == 0 for r in rem):
                res.append(tuple(cur))
            return
        i, j = pairs[p]
        hi = min(rem[i], rem[j])
        for v in range(hi + 1):
            rem[i] -= v
            rem[j] -= v
            cur.append(v)
            rec(p + 1, rem, cur)
            cur.pop()
            rem[i] += v
            rem[j] += v

    rec(0, list(nus), [])
    return pairs, res


def _canon(types, adj):
    """Canonical form of a labelled multigraph with stubs.

    types: tuple of (mu, nu) per vertex; adj: dict {(i,j): n} i<j.  Returns
    (key, perms) where perms lists every permutation p (new position -> old
    vertex) that realises the minimal key."""
    n = len(types)
    best = None
    perms = []
    for p in itertools.permutations(range(n)):
        t = tuple(types[p[i]] for i in range(n))
        a = tuple(adj.get((min(p[i], p[j]), max(p[i], p[j])), 0) for i in range(n) for j in range(i + 1, n))
        key = (t, a)
        if best is None or key < best:
            best = key
            perms = [p]
        elif key == best:
            perms.append(p)
    return best, perms


def enumerate_basis_graphs(level: int, max_factors: int = 4):
    """All basis functions of an MTP of the given level as canonical multigraph keys."""
    vtypes = []
    mu = 0
    while moment_level(mu, 0) <= level:
        nu = 0
        while moment_level(mu, nu) <= level:
            vtypes.append((mu, nu))
            nu += 1
        mu += 1
    found = set()
    for nv in range(1, max_factors + 1):
        for combo in itertools.combinations_with_replacement(vtypes, nv):
            if sum(moment_level(*t) for t in combo) > level:
                continue
            nus = [t[1] for t in combo]
            if nv == 1:
                if nus[0] == 0:
                    found.add(_canon(combo, {})[0])
                continue
            if sum(nus) % 2:
                continue
            pairs, mats = _row_sum_matrices(nus)
            for m in mats:
                adj = {pq: v for pq, v in zip(pairs, m) if v}
                found.add(_canon(combo, adj)[0])

    def sort_key(key):
        types, _ = key
        return (sum(moment_level(*t) for t in types), len(types), key)

    return sorted(found, key=sort_key)


# --------------------------------------------------------------------------- lowering

@dataclass
class _Node:
    key: tuple
    types: tuple
    adj: dict                # {(i,j): n}, canonical labelling
    stubs: tuple             # open indices per vertex
    split: tuple | None = None   # (left vertex ids, right vertex ids) in canonical labelling
    comps: dict = field(default_factory=dict)   # canonical component tuple -> temp moment id


class MTPBasisBuilder:
    def __init__(self):
        self.nodes: dict[tuple, _Node] = {}
        self.basic: dict[tuple, int] = {}          # (mu, ax, ay, az) -> temp id
        self.edges: list[list[int]] = []            # [a0, a1, mult, a3] temp ids
        self.nmom = 0
        self.order: list[int] = []                  # creation order of non-basic ids

    # -- helpers
    def _new_id(self):
        self.nmom += 1
        return self.nmom - 1

    @staticmethod
    def _adj_from_key(key):
        types, a = key
        n = len(types)
        adj = {}
        it = iter(a)
        for i in range(n):
            for j in range(i + 1, n):
                v = next(it)
                if v:
                    adj[(i, j)] = v
        return adj

    def _sub(self, types, adj, stubs, verts):
        """Canonical (key, perms) of the sub-network induced on verts; open indices are recorded
        through an extended vertex type (mu, nu, stub)."""
        idx = {v: i for i, v in enumerate(verts)}
        st = tuple((types[v][0], types[v][1], stubs_v) for v, stubs_v in ((v, self._stub_in(types, adj, v, verts)) for v in verts))
        sadj = {}
        for (i, j), n in adj.items():
            if i in idx and j in idx:
                a, b = idx[i], idx[j]
                sadj[(min(a, b), max(a, b))] = n
        return _canon(st, sadj)

    @staticmethod
    def _stub_in(types, adj, v, verts):
        inside = sum(n for (i, j), n in adj.items() if (i == v and j in verts) or (j == v and i in verts))
        return types[v][1] - inside

    @staticmethod
    def _ncomp(stubs):
        r = 1
        for s in stubs:
            r *= (s + 1) * (s + 2) // 2
        return r

    def _plan_cost(self, types, adj, verts, memo):
        """Cheapest binary contraction tree for the sub-network on verts.
        Cost = sum over not-yet-built intermediates of (#components x contraction fan-in)."""
        verts = tuple(sorted(verts))
        if verts in memo:
            return memo[verts]
        key, _ = self._sub(types, adj, None, verts)
        if len(verts) == 1 or key in self.nodes:
            memo[verts] = (0, None)
            return memo[verts]
        best = None
        vs = list(verts)
        first = vs[0]
        rest = vs[1:]
        for r in range(0, len(rest)):
            for extra in itertools.combinations(rest, r):
                left = (first,) + extra
                right = tuple(v for v in vs if v not in left)
                if not right:
                    continue
                cl, _ = self._plan_cost(types, adj, left, memo)
                cr, _ = self._plan_cost(types, adj, right, memo)
                stubs = [self._stub_in(types, adj, v, verts) for v in verts]
                fan = 1
                for (i, j), n in adj.items():
                    if (i in left and j in right) or (j in left and i in right):
                        fan *= (n + 1) * (n + 2) // 2
                cost = cl + cr + self._ncomp(stubs) * fan
                # tie-break towards balanced trees (shallower waves)
                cand = (cost, abs(len(left) - len(right)), left, right)
                if best is None or cand < best:
                    best = cand
        memo[verts] = (best[0], (best[2], best[3]))
        return memo[verts]

    def _ensure_node(self, types, adj, verts):
        """Create (recursively) the intermediate for the sub-network on verts.
        Returns (node, perm) with perm: canonical position -> vertex id in the caller's labelling."""
        verts = tuple(sorted(verts))
        key, perms = self._sub(types, adj, None, verts)
        perm = tuple(verts[p] for p in perms[0])
        if key in self.nodes:
            return self.nodes[key], perm
        ctypes_, _ = key
        cadj = self._adj_from_key(key)
        stubs = tuple(t[2] for t in ctypes_)
        node = _Node(key=key, types=tuple((t[0], t[1]) for t in ctypes_), adj=cadj, stubs=stubs)
        if len(verts) > 1:
            memo = {}
            _, split = self._plan_cost(types, adj, verts, memo)
            left, right = split
            inv = {v: i for i, v in enumerate(perm)}
            node.split = (tuple(inv[v] for v in left), tuple(inv[v] for v in right))
        self.nodes[key] = node
        return node, perm

    # -- components
    def _basic_id(self, mu, beta):
        k = (mu,) + tuple(beta)
        if k not in self.basic:
            self.basic[k] = self._new_id()
        return self.basic[k]

    def _automorphisms(self, node):
        n = len(node.types)
        et = tuple((node.types[i][0], node.types[i][1], node.stubs[i]) for i in range(n))
        _, perms = _canon(et, node.adj)
        return perms

    def component(self, node: _Node, comp: tuple) -> int:
        """Temp moment id of component ``comp`` (one symmetric multi-index per vertex, canonical order)."""
        n = len(node.types)
        if n == 1:
            return self._basic_id(node.types[0][0], comp[0])
        # merge symmetric duplicates: minimal image under the automorphism group
        if not hasattr(node, "_autos"):
            node._autos = self._automorphisms(node)
        comp = min(tuple(comp[p[i]] for i in range(n)) for p in node._autos)
        if comp in node.comps:
            return node.comps[comp]
        tid = self._new_id()
        node.comps[comp] = tid
        self.order.append(tid)
        left, right = node.split
        # the children, expressed in this node's labelling (vertex types keep full nu)
        lnode, lperm = self._ensure_node(node.types, node.adj, left)
        rnode, rperm = self._ensure_node(node.types, node.adj, right)
        cross = [((i, j) if i in left else (j, i), m) for (i, j), m in node.adj.items()
                 if (i in left) != (j in left)]
        cross = [(lr, m) for lr, m in cross]
        gamma_sets = [multi_indices(m) for _, m in cross]
        for gam in itertools.product(*gamma_sets):
            mult = 1
            add = {v: [0, 0, 0] for v in range(n)}
            for ((l, r), _), g in zip(cross, gam):
                mult *= multinomial(g)
                for a in range(3):
                    add[l][a] += g[a]
                    add[r][a] += g[a]
            lcomp = tuple(tuple(comp[v][a] + add[v][a] for a in range(3)) for v in lperm)
            rcomp = tuple(tuple(comp[v][a] + add[v][a] for a in range(3)) for v in rperm)
            a0 = self.component(lnode, lcomp)
            a1 = self.component(rnode, rcomp)
            self.edges.append([a0, a1, mult, tid])
        return tid

    def add_basis_function(self, key) -> int:
        types, _ = key
        adj = self._adj_from_key(key)
        node, perm = self._ensure_node(types, adj, tuple(range(len(types))))
        assert all(s == 0 for s in node.stubs)
        return self.component(node, tuple((0, 0, 0) for _ in types))


@dataclass
class MTPTables:
    level: int
    radial_funcs_count: int
    alpha_moments_count: int
    alpha_index_basic: list      # K x 4
    alpha_index_times: list      # T x 4
    alpha_moment_mapping: list   # A
    wave_sizes: list             # edges per dependency wave

    @property
    def alpha_index_basic_count(self):
        return len(self.alpha_index_basic)

    @property
    def alpha_index_times_count(self):
        return len(self.alpha_index_times)

    @property
    def alpha_scalar_moments(self):
        return len(self.alpha_moment_mapping)


@lru_cache(maxsize=None)
def build_mtp_tables(level: int, max_factors: int = 4) -> MTPTables:
    graphs = enumerate_basis_graphs(level, max_factors)
    b = MTPBasisBuilder()
    scalars = [b.add_basis_function(g) for g in graphs]

    # renumber: basics first, ordered (mu, nu, MLIP component order); then by wave, creation order
    def bkey(k):
        mu, ax, ay, az = k
        return (mu, ax + ay + az, -ax, -ay, -az)

    basics = sorted(b.basic, key=bkey)
    newid = {b.basic[k]: i for i, k in enumerate(basics)}
    wave = {tid: 0 for tid in newid}
    by_target = {}
    for e in b.edges:
        by_target.setdefault(e[3], []).append(e)

    def wave_of(tid):
        if tid in wave:
            return wave[tid]
        w = 1 + max(max(wave_of(e[0]), wave_of(e[1])) for e in by_target[tid])
        wave[tid] = w
        return w

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    created = {tid: i for i, tid in enumerate(b.order)}
    nonbasic = sorted(b.order, key=lambda t: (wave_of(t), created[t]))
    for tid in nonbasic:
        newid[tid] = len(newid)

    times = []
    for tid in nonbasic:
        merged = {}
        for a0, a1, mult, _ in by_target[tid]:
            p = (min(newid[a0], newid[a1]), max(newid[a0], newid[a1]))
            merged[p] = merged.get(p, 0) + mult
        for (a0, a1), mult in merged.items():
            times.append([a0, a1, mult, newid[tid]])
    nw = max([wave_of(t) for t in nonbasic], default=0)
    wave_sizes = [0] * nw
    for tid in nonbasic:
        wave_sizes[wave_of(tid) - 1] += len({(min(e[0], e[1]), max(e[0], e[1])) for e in by_target[tid]})
    mapping = sorted(newid[s] for s in scalars)
    assert len(set(mapping)) == len(mapping)
    R = 1 + max(k[0] for k in basics)
    return MTPTables(level=level, radial_funcs_count=R, alpha_moments_count=len(newid),
                     alpha_index_basic=[list(k) for k in basics], alpha_index_times=times,
                     alpha_moment_mapping=mapping, wave_sizes=wave_sizes)


def prepare_waves(alpha_index_times, alpha_index_basic_count):
    """Wave splitter with the semantics of the reference's block-parallel style
    (pair_mtps_kokkos.cpp:179-200), generalised to any depth: returns the list of
    wave sizes (the reference stores at most three and aborts on a fourth)."""
    sizes = []
    last_max_node = alpha_index_basic_count - 1
    last_max_edge = 0
    for i, e in enumerate(alpha_index_times):
        if e[0] > last_max_node or e[1] > last_max_node:
            sizes.append(i - last_max_edge)
            last_max_node = alpha_index_times[i - 1][3]
            last_max_edge = i
    sizes.append(len(alpha_index_times) - last_max_edge)
    return sizes


# MLIP level-8 template (public MLIP ``08.mtp``) -- known-answer table
LEVEL8_KAT = dict(
    radial_funcs_count=2, alpha_moments_count=18,
    alpha_index_basic=[[0, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0, 2, 0, 0], [0, 1, 1, 0],
                       [0, 1, 0, 1], [0, 0, 2, 0], [0, 0, 1, 1], [0, 0, 0, 2], [1, 0, 0, 0]],
    alpha_index_times=[[0, 0, 1, 11], [1, 1, 1, 12], [2, 2, 1, 12], [3, 3, 1, 12], [4, 4, 1, 13],
                       [5, 5, 2, 13], [6, 6, 2, 13], [7, 7, 1, 13], [8, 8, 2, 13], [9, 9, 1, 13],
                       [0, 10, 1, 14], [0, 11, 1, 15], [0, 12, 1, 16], [0, 15, 1, 17]],
    alpha_moment_mapping=[0, 10, 11, 12, 13, 14, 15, 16, 17],
)
