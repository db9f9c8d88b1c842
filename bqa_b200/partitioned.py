"""Node-partitioned engine: one process per GPU, boundary messages exchanged once per BP sweep.

The reference is single-process; this is the B200 scaling axis named by BASELINE.json:
qubits are partitioned across the GPUs of one box, the owner of a node computes all its outgoing messages, and a
directed message n -> m with owner(n) != owner(m) is a *boundary* message.

Per BP sweep (reference state.py:105-121):   sweep kernels of the owned degree classes
                                             -> grouped send/recv of the boundary messages into the peers' halo slots
                                             -> max all-reduce of the two residual scalars (get_dist, backends.py:492-495)
Per annealing step (state.py:230-247):       extended messages of the cut edges are exchanged so that both owners hold
                                             the pair (m_f, m_b); both run the same canonicalizer kernel on the same bits
                                             (identical gauge, no canonicalizer traffic); the column maxima behind the
                                             global bond-dimension decision (backends.py:297-303) are max all-reduced.

Every rank runs the unmodified single-GPU kernels on a *local context*: its owned nodes grouped by degree and the
undirected edges touching them, renumbered so that the reference's slot convention (forward edge e at slot e,
backward at e + L, lambda at slot mod L; config_canonicalization.py:189-205, state.py:175-183) holds locally.
Collectives go through torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
import logging
import os
from dataclasses import dataclass, replace

import numpy as np
import torch
import torch.distributed as dist

from .config import Context, Layout
from .engine import Engine, _np, _on_device

log = logging.getLogger(__name__)


# ---------------------------------------------------------------------------------------------------
# partitioner: balanced multi-seed region growing (no METIS in the image; SURVEY.md section 8e measured
# 0.17 / 0.25 / 0.29 cut-edge fractions at P = 2 / 4 / 8 on the 100k-qubit 3-regular instance for this scheme)
# ---------------------------------------------------------------------------------------------------
def _csr(n: int, edges: np.ndarray):
    src = np.concatenate([edges[:, 0], edges[:, 1]])
    dst = np.concatenate([edges[:, 1], edges[:, 0]])
    order = np.argsort(src, kind="stable")
    src, dst = src[order], dst[order]
    indptr = np.zeros(n + 1, np.int64)
    np.add.at(indptr, src + 1, 1)
    return np.cumsum(indptr), dst


def _neighbours(indptr, indices, nodes):
    if nodes.size == 0:
        return nodes
    starts, ends = indptr[nodes], indptr[nodes + 1]
    lens = ends - starts
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64)
    offs = np.repeat(starts - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(total)
    return indices[offs]


def _bfs_far(indptr, indices, n, start):
    """Last node reached by a BFS from `start` (a far-away node of its component)."""
    seen = np.zeros(n, bool)
    seen[start] = True
    frontier = np.array([start])
    last = start
    while frontier.size:
        last = int(frontier[-1])
        nb = np.unique(_neighbours(indptr, indices, frontier))
        nb = nb[~seen[nb]]
        seen[nb] = True
        frontier = nb
    return last


def partition_nodes(n_nodes: int, edges: np.ndarray, n_parts: int, seed: int = 0, refine_passes: int = 4) -> np.ndarray:
    """part[node] in [0, n_parts): balanced (sizes differ by at most one) connected-ish regions grown breadth
    first from far-apart seeds, followed by greedy boundary refinement.  Deterministic in (graph, n_parts, seed)."""
    if n_parts <= 1:
        return np.zeros(n_nodes, np.int32)
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    indptr, indices = _csr(n_nodes, edges)
    rng = np.random.default_rng(seed)
    cap = np.full(n_parts, n_nodes // n_parts, np.int64)
    cap[: n_nodes % n_parts] += 1
    part = np.full(n_nodes, -1, np.int32)
    seeds = [_bfs_far(indptr, indices, n_nodes, int(rng.integers(n_nodes)))]
    dist_to_seeds = None
    for _ in range(1, n_parts):                              # farthest-point seeding by hop distance
        d = np.full(n_nodes, np.iinfo(np.int32).max, np.int64)
        d[seeds[-1]] = 0
        frontier = np.array([seeds[-1]])
        level = 0
        while frontier.size:
            level += 1
            nb = np.unique(_neighbours(indptr, indices, frontier))
            nb = nb[d[nb] > level]
            d[nb] = level
            frontier = nb
        dist_to_seeds = d if dist_to_seeds is None else np.minimum(dist_to_seeds, d)
        cand = np.where(dist_to_seeds < np.iinfo(np.int32).max, dist_to_seeds, -1)
        cand[seeds] = -1
        seeds.append(int(np.argmax(cand)))
    sizes = np.zeros(n_parts, np.int64)
    frontiers = []
    for r, s in enumerate(seeds):
        if part[s] >= 0:                                     # duplicate seed (tiny graphs): take any free node
            free = np.flatnonzero(part < 0)
            s = int(free[0])
        part[s] = r
        sizes[r] = 1
        frontiers.append(np.array([s]))
    while True:
        grew = False
        for r in np.argsort(sizes, kind="stable"):           # smallest region grows first
            room = int(cap[r] - sizes[r])
            if room <= 0 or frontiers[r].size == 0:
                continue
            nb = np.unique(_neighbours(indptr, indices, frontiers[r]))
            nb = nb[part[nb] < 0]
            if nb.size > room:
                nb = nb[:room]
            if nb.size:
                part[nb] = r
                sizes[r] += nb.size
                grew = True
            frontiers[r] = nb
        if not grew:
            break
    free = np.flatnonzero(part < 0)                          # enclosed pockets / other components
    if free.size:                                            # fill the remaining capacity; refinement repairs the cut
        room = np.repeat(np.arange(n_parts), np.maximum(cap - sizes, 0))
        part[free] = room[: free.size].astype(np.int32)
    for _ in range(refine_passes):
        if not _refine(part, indptr, indices, n_parts, cap):
            break
    return part


def _refine(part, indptr, indices, n_parts, cap) -> bool:
    """One greedy pass: pairs of boundary nodes with positive gain swap sides (keeps the sizes exactly)."""
    n = part.shape[0]
    deg = np.diff(indptr)
    src = np.repeat(np.arange(n), deg)
    same = np.zeros(n, np.int64)
    np.add.at(same, src, (part[src] == part[indices]).astype(np.int64))
    moved = False
    # best foreign part of every boundary node
    foreign = part[indices] != part[src]
    if not foreign.any():
        return False
    bs, bp = src[foreign], part[indices][foreign]
    key = bs * n_parts + bp
    uniq, cnt = np.unique(key, return_counts=True)
    node, tgt = uniq // n_parts, (uniq % n_parts).astype(np.int32)
    gain = cnt - same[node]
    order = np.argsort(-gain, kind="stable")
    node, tgt, gain = node[order], tgt[order], gain[order]
    first = np.unique(node, return_index=True)[1]
    node, tgt, gain = node[first], tgt[first], gain[first]
    keep = gain > 0
    node, tgt, gain = node[keep], tgt[keep], gain[keep]
    if node.size == 0:
        return False
    # match movers r -> q with movers q -> r, best gains first; nodes adjacent to an already moved node are skipped
    locked = np.zeros(n, bool)
    order = np.argsort(-gain, kind="stable")
    buckets = {}
    for i in order:
        v, q, r = int(node[i]), int(tgt[i]), int(part[node[i]])
        if locked[v]:
            continue
        partner_list = buckets.get((q, r))
        if partner_list:
            u = partner_list.pop()
            if locked[u]:
                continue
            part[v], part[u] = q, r
            for w in (v, u):
                locked[w] = True
                locked[indices[indptr[w]:indptr[w + 1]]] = True
            moved = True
        else:
            buckets.setdefault((r, q), []).append(v)
    return moved


def cut_fraction(part: np.ndarray, edges: np.ndarray) -> float:
    edges = np.asarray(edges).reshape(-1, 2)
    return float(np.mean(part[edges[:, 0]] != part[edges[:, 1]])) if edges.size else 0.0


# ---------------------------------------------------------------------------------------------------
# local context + exchange plan
# ---------------------------------------------------------------------------------------------------
@dataclass
class ExchangePlan:
    rank: int
    world: int
    owned: np.ndarray                 # global ids of the owned nodes, in local id order
    send_slots: dict                  # peer -> local slots whose freshly computed messages go to that peer
    recv_slots: dict                  # peer -> local halo slots filled by that peer (same global-position order)
    n_local_edges: int
    n_cut_edges: int
    max_local_edges: int = 0          # over all ranks (symmetric buffers are sized for it)
    n_boundary_nodes: dict = None     # degree -> owned nodes with a remote out-edge (they come first in the class)
    canon_owned: np.ndarray = None    # local edge indices this rank canonicalizes (inner edges + its share of the cut edges)
    canon_remote: np.ndarray = None   # per local edge: -1 or (peer << 27 | the peer's index of the edge), for owned cut edges
    local_edges_all: list = None      # local edge count of every rank


def build_local_context(ctx: Context, part: np.ndarray, rank: int, world: int):
    """Local Context (owned nodes, edges touching them, reference slot convention kept) and the exchange plan."""
    E = np.asarray(ctx.edges, np.int64).reshape(-1, 2)
    L = E.shape[0]
    pl, pr = part[E[:, 0]], part[E[:, 1]]
    local_edge = (pl == rank) | (pr == rank)
    g2l = np.full(L, -1, np.int64)
    le = np.flatnonzero(local_edge)
    g2l[le] = np.arange(le.size)
    Lr = int(le.size)

    def slot(pos):                                           # global message position -> local slot
        pos = np.asarray(pos, np.int64)
        return g2l[pos % L] + (pos // L) * Lr

    # slot numbering of every rank (for the peer-memory path: a boundary message is stored straight into the
    # halo slot of the rank that owns its receiver)
    g2l_all, L_all = [], []
    for q in range(world):
        mq = (pl == q) | (pr == q)
        gq = np.full(L, -1, np.int64)
        gq[mq] = np.arange(int(mq.sum()))
        g2l_all.append(gq)
        L_all.append(int(mq.sum()))

    def remote(pos):                                         # global position of an outgoing message -> remote code
        pos = np.asarray(pos, np.int64)
        e, back = pos % L, pos // L
        recv_owner = np.where(back == 0, pr[e], pl[e])       # forward = lhs -> rhs, backward = rhs -> lhs
        code = np.full(pos.shape, -1, np.int64)
        for q in range(world):
            if q == rank:
                continue
            m = recv_owner == q
            code[m] = (q << 27) | (g2l_all[q][e[m]] + back[m] * L_all[q])
        return code

    owned = np.flatnonzero(part == rank)
    gid2lid = np.full(part.shape[0], -1, np.int64)
    gid2lid[owned] = np.arange(owned.size)
    layouts = {}
    n_boundary = {}
    for d, lay in ctx.degree_to_layout.items():
        ids = _np(lay.node_ids).astype(np.int64)
        m = part[ids] == rank
        if not m.any():
            continue
        d = int(d)
        # owned nodes of the class, those with a remote out-edge FIRST (stable): the single-launch BP run sweeps them
        # first and sends its "halo ready" line while the interior nodes are still being swept
        rem = remote(_np(lay.output_msgs_position).reshape(d, -1)[:, m]) if d > 0 else np.zeros((0, int(m.sum())), np.int64)
        is_b = (rem >= 0).any(axis=0) if d > 0 else np.zeros(int(m.sum()), bool)
        order = np.argsort(~is_b, kind="stable")
        n_boundary[d] = int(is_b.sum())
        sel = lambda a, order=order, m=m, d=d: _np(a).reshape(d, -1)[:, m][:, order]
        sel1 = lambda a, order=order, m=m: np.asarray(a)[m][order]
        layouts[d] = Layout(
            node_ids=gid2lid[sel1(ids)],
            input_msgs_position=slot(sel(lay.input_msgs_position)),
            output_msgs_position=slot(sel(lay.output_msgs_position)),
            lmbds_position=g2l[np.asarray(sel(lay.lmbds_position), np.int64)],
            node_ampls=sel1(np.real(_np(lay.node_ampls))), edge_ampls=np.real(sel(lay.edge_ampls)),
            remote_msgs_position=remote(sel(lay.output_msgs_position)))
    # boundary messages: forward position e is lhs -> rhs, backward position e + L is rhs -> lhs
    cut = np.flatnonzero(local_edge & (pl != pr))
    send, recv = {}, {}
    for e in cut:
        if pl[e] == rank:                                    # lhs owned: send forward, receive backward
            peer, s_pos, r_pos = int(pr[e]), e, e + L
        else:
            peer, s_pos, r_pos = int(pl[e]), e + L, e
        send.setdefault(peer, []).append(s_pos)
        recv.setdefault(peer, []).append(r_pos)
    assert max(L_all) < (1 << 26), "too many local edges for the 27-bit remote slot code"
    # one owner per cut edge (alternating by edge parity between the two endpoint ranks): it canonicalizes the edge and
    # stores the result into the other rank as well (bqa_b200_canonicalize_p2p)
    el = E[le]
    ple, pre = part[el[:, 0]], part[el[:, 1]]
    is_cut = ple != pre
    owner = np.where(le % 2 == 0, ple, pre)
    mine = ~is_cut | (owner == rank)
    other = np.where(ple == rank, pre, ple)
    canon_remote = np.full(Lr, -1, np.int64)
    sel_c = is_cut & (owner == rank)
    for q in range(world):
        mq = sel_c & (other == q)
        canon_remote[mq] = (q << 27) | g2l_all[q][le[mq]]
    plan = ExchangePlan(rank=rank, world=world, owned=owned, max_local_edges=max(L_all),
                        send_slots={q: slot(np.sort(np.array(v))) for q, v in send.items()},
                        recv_slots={q: slot(np.sort(np.array(v))) for q, v in recv.items()},
                        n_local_edges=Lr, n_cut_edges=int(cut.size), n_boundary_nodes=n_boundary,
                        canon_owned=np.flatnonzero(mine).astype(np.int32), canon_remote=canon_remote.astype(np.int32),
                        local_edges_all=L_all)
    local = replace(ctx, nodes_number=int(owned.size), edges_number=2 * Lr, degree_to_layout=layouts,
                    edges=E[le], couplings=np.asarray(ctx.couplings)[le], fields=np.asarray(ctx.fields)[owned],
                    node_degree=None, node_slot=None)
    return local, plan


# ---------------------------------------------------------------------------------------------------
# engine
# ---------------------------------------------------------------------------------------------------
class PartitionedEngine(Engine):
    """Same interface as Engine; ``bloch_vectors`` / ``measure`` return global results on every rank.

    Two transports for the boundary messages:
      * peer memory (default on CUDA + NCCL): message / extended-message / control buffers live in symmetric
        memory, the sweep kernels store boundary messages straight into the owner's halo slots over NVLink and
        ``bqa_b200_sweep_sync`` pushes the residual maxima and runs the barrier -- two kernel launches per sweep,
        no pack / unpack, no collective call in the sweep loop;
      * ``torch.distributed`` P2P ops + all-reduce (``p2p=False``; the only transport under gloo / in the CPU tests).
    """

    def __init__(self, context, precision=None, device=None, group=None, part=None, partition_seed: int = 0,
                 p2p: bool | None = None, _testing_lib=None):
        if not dist.is_initialized():
            raise RuntimeError("PartitionedEngine needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.global_ctx = context
        self.N_global = int(context.nodes_number)
        E = np.asarray(context.edges, np.int64).reshape(-1, 2)
        self.part = np.asarray(part, np.int32) if part is not None else \
            partition_nodes(self.N_global, E, self.world, seed=partition_seed)
        local, plan = build_local_context(context, self.part, self.rank, self.world)
        self.plan = plan
        if p2p is None:
            p2p = (_testing_lib is None and dist.get_backend(group) == "nccl" and self.world <= 8
                   and os.environ.get("BQA_B200_P2P", "1") != "0")
        self.p2p = bool(p2p) and self.world > 1
        if self.p2p:                                          # every pair of GPUs must be peer-accessible; all ranks agree
            dev_idx = torch.device(device).index if device is not None else torch.cuda.current_device()
            ok = all(d == dev_idx or torch.cuda.can_device_access_peer(dev_idx, d) for d in range(torch.cuda.device_count()))
            flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", dev_idx))
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if not int(flag.item()):
                log.warning("peer access between the GPUs is not available: boundary messages go through torch.distributed")
                self.p2p = False
        self._symm = {}                                       # tag -> (tensor, handle)
        self._seq = 0
        super().__init__(local, precision=precision, device=device, _testing_lib=_testing_lib)
        self.rng = np.random.default_rng(int(context.seed))              # same stream on every rank
        self._multiclass = False          # the table-driven launches have no halo stores / no exchange between sweeps
        dev = self.dev
        self._peers = sorted(plan.send_slots)
        self._send_idx = {q: torch.from_numpy(plan.send_slots[q]).to(dev) for q in self._peers}
        self._recv_idx = {q: torch.from_numpy(plan.recv_slots[q]).to(dev) for q in self._peers}
        self._owned_dev = torch.from_numpy(plan.owned).to(dev)
        self.comm_bytes = 0
        # every rank must run BP the same way (the in-kernel barriers of the single-launch run and the per-sweep
        # launches stop after different numbers of barriers): single launch only if it is possible everywhere
        active = [c.degree for c in self.classes if c.degree > 0 and c.B > 0]
        ok = len(active) == 1 and self._single_launch_ok and self.p2p
        ok = ok and all(c.B >= 4 for c in self.classes if c.degree > 0 and c.B > 0)   # groups of 4 nodes (bqa_fast_d3D4.cu)
        deg = active[0] if ok else -1
        flag = torch.tensor([1 if ok else 0, deg, -deg], device=dev)      # MIN over ranks: all ok, min degree, -max degree
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        self._single_launch_ok = bool(int(flag[0])) and int(flag[1]) == -int(flag[2])   # same single class everywhere
        if self.p2p:
            self._ensure_edge_buffers(self.Dmax)              # symmetric allocations are collective: do them all now
            # 64 bytes of per-peer barrier flags (bqa_b200_sweep_sync) + the handshake lines of the single-launch BP run
            self._flags = self._alloc_shared(64 + 8 * 16 * 8, torch.uint8, "flags")
            ptrs = lambda tag: (C.c_void_p * 8)(*([int(x) for x in self._symm[tag][1].buffer_ptrs] + [0] * (8 - self.world)))
            self._peer_ptrs = {("msgs", 0): ptrs("msgs0"), ("msgs", 1): ptrs("msgs1"), ("msgs", 2): ptrs("msgs2"),
                               ("ext", 0): ptrs("ext"), "canon": ptrs("canon"), "lmbds": ptrs("lmbds"),
                               "ctrl": ptrs("ctrl"), "flags": ptrs("flags")}
            self._canon_owned = torch.from_numpy(plan.canon_owned).to(dev)
            self._canon_remote = torch.from_numpy(plan.canon_remote).to(dev)
            self._canon_peer_L = (C.c_longlong * 8)(*([int(x) for x in plan.local_edges_all] + [0] * (8 - self.world)))
            dist.barrier(group=self.group)
        log.info(f"rank {self.rank}/{self.world}: {plan.owned.size} nodes, {plan.n_local_edges} local edges, "
                 f"{plan.n_cut_edges} cut, transport {'peer memory' if self.p2p else 'torch.distributed'}")

    def _n_msg_buffers(self) -> int:
        return 3 if self.p2p else 2

    # -- symmetric memory ------------------------------------------------------------------------------
    def _alloc_shared(self, numel: int, dtype, tag: str) -> torch.Tensor:
        if not self.p2p:
            return super()._alloc_shared(numel, dtype, tag)
        import torch.distributed._symmetric_memory as symm_mem
        esize = torch.empty(0, dtype=dtype).element_size()
        if tag.startswith("msgs"):                            # same size on every rank: the largest local slot count
            numel = 2 * self.plan.max_local_edges * self.Dmax * self.Dmax
        elif tag in ("ext", "canon"):
            numel = 2 * self.plan.max_local_edges * 4 * self.Dmax * self.Dmax
        elif tag == "lmbds":
            numel = self.plan.max_local_edges * 2 * self.Dmax
        nbytes = (numel * esize + 511) // 512 * 512
        raw = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.dev)
        raw.zero_()
        pg = self.group if self.group is not None else dist.group.WORLD
        handle = symm_mem.rendezvous(raw, pg)
        self._symm[tag] = (raw, handle)
        return raw[: numel * esize].view(dtype)

    def _peer_targets(self, what: str, parity: int = 0):
        return self._peer_ptrs[(what, parity)] if self.p2p else None

    def _sync(self, it: int) -> None:
        """Residual push (it >= 0) + barrier over peer memory: one tiny kernel."""
        self._seq += 1
        self.lib.sweep_sync(self.prec, self.rank, self.world, self._peer_ptrs["ctrl"], it, self._peer_ptrs["flags"],
                            self._seq, self._status.data_ptr(), self._stream())

    def _bp_run_peers(self):
        if not self.p2p:
            return None                # the torch.distributed transport exchanges between sweeps: no single launch
        pp = self._peer_ptrs
        active = [c for c in self.classes if c.degree > 0 and c.B > 0]
        n_boundary = int((self.plan.n_boundary_nodes or {}).get(active[0].degree, active[0].B)) if active else 0
        return (self.rank, self.world, pp[("msgs", 0)], pp[("msgs", 1)], pp["ctrl"], pp["flags"], self._seq,
                pp[("msgs", 2)], n_boundary)

    def _bp_run_done(self, sweeps: int) -> None:
        self._seq += self.max_iters + 1    # sequence numbers a run may have used (one per executed sweep, <= max + 1)

    def _before_bp(self) -> None:
        # per-sweep launches: nobody pushes residuals of the new run before everybody has reset its control block.
        # (The single-launch run writes nothing into a peer's control block -- handshake lines only: no barrier.)
        if self.p2p and not self._single_launch_applies():
            self._sync(-1)

    # -- boundary exchange of a (slots, elems) complex array held flat in `buf` ------------------------
    def _exchange(self, buf: torch.Tensor, elems: int) -> None:
        if not self._peers:
            return
        view = buf[: self.E2 * elems].view(self.E2, elems)
        ops, recvs = [], []
        for q in self._peers:
            send = view.index_select(0, self._send_idx[q]).contiguous()
            recv = torch.empty((self._recv_idx[q].numel(), elems), dtype=view.dtype, device=view.device)
            # complex tensors travel as real pairs (gloo has no complex support)
            ops.append(dist.P2POp(dist.isend, torch.view_as_real(send), self._global_rank(q), self.group))
            ops.append(dist.P2POp(dist.irecv, torch.view_as_real(recv), self._global_rank(q), self.group))
            recvs.append((q, recv))
            self.comm_bytes += send.numel() * send.element_size()
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for q, recv in recvs:
            view.index_copy_(0, self._recv_idx[q], recv)

    def _global_rank(self, q: int) -> int:
        return dist.get_global_rank(self.group, q) if self.group is not None else q

    # -- hooks of the single-GPU engine ------------------------------------------------------------------
    def _after_sweep(self, it: int, nxt: torch.Tensor) -> None:
        if self.p2p:
            self._sync(it)
            return
        self._exchange(nxt, self.D * self.D)
        dist.all_reduce(self._resid[2 * it: 2 * it + 2], op=dist.ReduceOp.MAX, group=self.group)

    def _exchange_ext(self) -> None:
        if self.p2p:
            self._sync(-1)
            return
        self._exchange(self._ext, 4 * self.D * self.D)

    def _canonicalize(self, D: int, st: int) -> None:
        """One owner per cut edge on the peer-memory path (the owner stores the result into the other rank too); the max
        all-reduce of the column maxima that follows on every rank's stream orders those stores before the apply kernels."""
        if not self.p2p or os.environ.get("BQA_B200_CANON_SINGLE_OWNER", "1") == "0":
            return super()._canonicalize(D, st)
        pp = self._peer_ptrs
        self.lib.canonicalize_p2p(self.prec, D, self.L, self._ext.data_ptr(), self._canon.data_ptr(), self._lmbds.data_ptr(),
                                  self._colmax.data_ptr(), self.pinv_eps, min(2 * D, self.Dmax), self._canon_owned.numel(),
                                  self._canon_owned.data_ptr(), self._canon_remote.data_ptr(), pp["canon"], pp["lmbds"],
                                  self._canon_peer_L, st)

    def _reduce_colmax(self, colmax: torch.Tensor) -> None:
        dist.all_reduce(colmax, op=dist.ReduceOp.MAX, group=self.group)

    def _after_update(self) -> None:
        # halo slots of the re-initialised messages: lambdas of cut edges are held by both owners, so every slot is
        # filled locally (state.py:56-57) instead of being exchanged
        self.lib.gauge_msgs(self.prec, self._lmbd_stride // 2, self.D, self.L, self._lmbds.data_ptr(),
                            self.msgs_buffer.data_ptr(), self._stream())

    # -- results ---------------------------------------------------------------------------------------
    # checkpoint hooks (core.save_checkpoint / load_checkpoint): one file per rank holding the rank's local state
    # (owned nodes, halo slots, lambdas of its edges); the partition is deterministic, so a resumed run rebuilds the
    # same local contexts
    def checkpoint_file(self, path: str) -> str:
        root, ext = os.path.splitext(path)
        return f"{root}.rank{self.rank}of{self.world}{ext}"

    def checkpoint_barrier(self) -> None:
        dist.barrier(group=self.group)

    def checkpoint_agree(self, value: int, what: str) -> int:
        t = torch.tensor([int(value), -int(value)], dtype=torch.int64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        lo, hi = int(t[0]), -int(t[1])
        if lo != hi:
            raise RuntimeError(f"the ranks disagree on the checkpoint's {what} ({lo} .. {hi}): the per-rank files "
                               "are not from the same save; delete them to start over")
        return lo

    def _gather_rows(self, local: torch.Tensor, width: int) -> torch.Tensor:
        """(N_global, width) array assembled from every rank's owned rows (all ranks get it)."""
        full = torch.zeros((self.N_global, width), dtype=local.dtype, device=local.device)
        full.index_copy_(0, self._owned_dev, local.view(-1, width))
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
        return full

    @_on_device
    def bloch_vectors(self) -> np.ndarray:
        self._compute_bloch()
        full = self._gather_rows(self._bloch, 4)
        log.info("Density matrices have been computed")
        return full.cpu().numpy()[:, :3].astype(np.float64)

    @_on_device
    def measure(self) -> list:
        """reference state.py:250-312 with the candidate search and the projections done by the owners."""
        st = self._stream()
        outcomes = torch.zeros(self.N, dtype=torch.int32, device=self.dev)
        owned = self.plan.owned
        while True:
            self._compute_bloch()
            self.lib.argmax_unmeasured(self.prec, self.N, self._bloch.data_ptr(), outcomes.data_ptr(),
                                       self._argmax_i.data_ptr(), self._argmax_p.data_ptr(), st)
            mine = self._to_host(self._argmax)
            node, left = (int(v) for v in mine[:8].view(np.int32))
            p0 = float(mine[8:8 + self._argmax_p.element_size()].view(self.np_rdtype)[0])
            cand = torch.tensor([abs(2.0 * p0 - 1.0) if left else -1.0, float(owned[node]) if left else -1.0, p0,
                                 float(left)], dtype=torch.float64, device=self.dev)
            allc = [torch.zeros_like(cand) for _ in range(self.world)]
            dist.all_gather(allc, cand, group=self.group)
            allc = torch.stack(allc).cpu().numpy()
            if allc[:, 3].sum() == 0:
                break
            # first node in id order maximising |2 p0 - 1| among the unmeasured (state.py:301); the key is
            # compared in the working precision, like the single-GPU kernel does
            keys = allc[:, 0].astype(self.np_rdtype)
            best = max(range(self.world), key=lambda r: (keys[r], -allc[r, 1]))
            gnode, p0 = int(allc[best, 1]), float(allc[best, 2])
            u = self.rng.uniform(0.0, 1.0)
            bit = 0 if p0 > u else 1
            if best == self.rank:
                c = self.classes[int(self._node_class[node])]
                self.lib.project_node(self.prec, c.degree, self.D, c.T[c.cur].data_ptr(), int(self._node_slot[node]), bit, st)
                outcomes[node] = 1 - 2 * bit
            self.run_bp()
            self._compute_bloch()
            for c in self.classes:
                self.lib.threshold_project(self.prec, c.degree, self.D, c.B, c.T[c.cur].data_ptr(), c.node_ids.data_ptr(),
                                           self._bloch.data_ptr(), outcomes.data_ptr(), self.threshold,
                                           self._nproj.data_ptr(), st)
            self.run_bp()
        full = self._gather_rows(outcomes.to(torch.float64), 1)
        return [int(v) for v in full.cpu().numpy().reshape(-1)]
