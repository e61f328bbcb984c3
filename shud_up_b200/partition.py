"""Domain decomposition of a SHUD mesh for one-process-per-GPU runs.

Owner-computes: a partition owns a set of cells (and the reaches / segments / lakes attached to them);
every edge flux is evaluated by the owner of the cell on its own side, so the only data that crosses a cut
is the state (Ysurf, Ygw) of the cells on the far side of cut edges - the HALO cells.  Nothing is summed
across partitions, so ydot of an owned cell is bit-identical to the single-GPU result.

    extract(mesh, owned_mask)      local mesh (owned cells renumbered 1..Ne, halo cells Ne+1..Ne+Nhalo)
    HaloExchange(local, comm)      per-RHS exchange: pack -> all_to_all -> halo buffers

Restrictions of this version (checked, NotImplementedError otherwise): a reach, its segments' cells and its
downstream reach live in one partition; a lake and its bank cells live in one partition; halo cells are
plain land cells (no head BC).  Host-side set-up only - no RHS arithmetic here.
"""
import numpy as np

CELL_SKIP = ("ele_nabr", "ele_lakenabr")
EDGE_KEYS = ("ele_edge", "ele_Dist2Nabor", "ele_Dist2Edge", "ele_avgRough")
HALO_KEYS = ("z_surf", "z_bottom", "AquiferDepth", "macD", "macKsatH", "geo_vAreaF", "KsatH")
CELL_DYN = ("qEleNetPrep", "qPotEvap", "qPotTran", "t_lai", "fu_Surf", "fu_Sub", "qElePrep", "qEleE_IC_in",
            "ele_yBC", "ele_QBC", "ele_u_satn")


def extract(mesh, owned_mask, gid=None, part_of_cell=None, keep_full_halo=False):
    """Cut the partition `owned_mask` (bool [Ne]) out of `mesh` (snapshot-named dict, 1-based indices).
    gid: global id of every cell of `mesh` (default: its index); carried along as own_gid / halo_gid so that
    partitions cut from different pieces of one global mesh agree on the order of exchanged cells.
    part_of_cell: owner rank of every cell; halo cells are then ordered by (owner, global id), which is the
    order the peers' messages arrive in, so the receive buffer IS the halo state array (no unpack)."""
    Ne, Nr, Ns, Nl = (int(np.asarray(mesh[k]).reshape(-1)[0]) for k in ("Ne", "Nr", "Ns", "Nl"))
    owned_mask = np.asarray(owned_mask, dtype=bool)
    gid = np.arange(Ne, dtype=np.int64) if gid is None else np.asarray(gid, dtype=np.int64)
    own = np.nonzero(owned_mask)[0]                     # ascending reference id
    nown = own.size
    nabr = np.asarray(mesh["ele_nabr"]).reshape(3, Ne)
    lnab = np.asarray(mesh["ele_lakenabr"]).reshape(3, Ne)
    nb_own = nabr[:, own]                               # [3][nown], 1-based, 0 = none
    has = nb_own > 0
    nb0 = np.where(has, nb_own - 1, 0)
    is_halo_ref = has & ~owned_mask[nb0]
    halo = np.unique(nb0[is_halo_ref])                  # reference ids of halo cells, ascending
    if part_of_cell is not None:
        halo = halo[np.lexsort((gid[halo], np.asarray(part_of_cell)[halo]))]   # by (owner, global id)
    else:
        halo = halo[np.argsort(gid[halo], kind="stable")]                      # by global id
    new_id = np.zeros(Ne, dtype=np.int64)               # 1-based local id of every cell that survives
    new_id[own] = np.arange(1, nown + 1)
    new_id[halo] = nown + np.arange(1, halo.size + 1)
    if np.any(np.asarray(mesh["ele_iBC"])[halo] > 0) or np.any(np.asarray(mesh["ele_iLake"])[halo] > 0):
        raise NotImplementedError("halo cells with a head BC or inside a lake are not supported")
    if np.any(lnab[:, own][is_halo_ref] > 0):
        raise NotImplementedError("a lake bank is cut by the partition")
    loc = {}
    for k, v in mesh.items():
        v = np.asarray(v)
        if k in CELL_SKIP:
            continue
        if k in EDGE_KEYS:
            loc[k] = np.ascontiguousarray(v.reshape(3, Ne)[:, own]).ravel()
        elif k.startswith("ele_") and v.ndim == 1 and v.shape[0] == Ne:
            loc[k] = v[own]
        elif k in CELL_DYN and v.shape[0] == Ne:
            loc[k] = v[own]
    loc["ele_nabr"] = np.where(has, new_id[nb0], 0).astype(np.int32).ravel()
    loc["ele_lakenabr"] = np.ascontiguousarray(lnab[:, own]).astype(np.int32).ravel()
    for k in HALO_KEYS:
        loc["halo_" + k] = np.asarray(mesh["ele_" + k])[halo]
    loc["own_gid"], loc["halo_gid"] = gid[own], gid[halo]
    if keep_full_halo:
        loc["_halo_ref"] = halo
    # ---- reaches and segments: owned by the partition of their segments' cells ----
    seg_e = np.asarray(mesh["seg_iEle"]) - 1
    seg_r = np.asarray(mesh["seg_iRiv"]) - 1
    seg_owned = owned_mask[seg_e] if Ns else np.zeros(0, dtype=bool)
    riv_any = np.zeros(Nr, dtype=bool); riv_all = np.ones(Nr, dtype=bool)
    if Ns:
        np.logical_or.at(riv_any, seg_r, seg_owned)
        np.logical_and.at(riv_all, seg_r, seg_owned)
    if np.any(riv_any & ~riv_all):
        raise NotImplementedError("a reach has segments in two partitions")
    if Nr:
        # a river tree is owned as a whole (reaches without segments follow the tree they belong to)
        comp = _reach_components(mesh["riv_down"])
        c_any = np.zeros(Nr, dtype=bool); c_all = np.ones(Nr, dtype=bool)
        has_seg = np.zeros(Nr, dtype=bool)
        if Ns:
            has_seg[seg_r] = True
        np.logical_or.at(c_any, comp, riv_any)
        np.logical_and.at(c_all, comp[has_seg], riv_any[has_seg])
        if np.any(c_any & ~c_all):
            raise NotImplementedError("a reach and its downstream reach are in different partitions")
        riv_any = c_any[comp]
        # trees without any segment are owned by the partition that owns cell 0 of the mesh
        orphan = ~has_seg
        c_has = np.zeros(Nr, dtype=bool)
        np.logical_or.at(c_has, comp, has_seg)
        riv_any = riv_any | (~c_has[comp] & bool(owned_mask[0]))
    rown = np.nonzero(riv_any)[0]
    down = np.asarray(mesh["riv_down"])
    dn = down[rown]
    if np.any((dn > 0) & ~riv_any[np.maximum(dn - 1, 0)]):
        raise NotImplementedError("a reach and its downstream reach are in different partitions")
    rnew = np.zeros(Nr, dtype=np.int64); rnew[rown] = np.arange(1, rown.size + 1)
    for k, v in mesh.items():
        v = np.asarray(v)
        if k.startswith("riv_") and v.ndim == 1 and v.shape[0] == Nr:
            loc[k] = v[rown]
    loc["riv_down"] = np.where(dn > 0, rnew[np.maximum(dn - 1, 0)], dn).astype(np.int32)
    sown = np.nonzero(seg_owned)[0]
    loc["seg_iEle"] = new_id[seg_e[sown]].astype(np.int32)
    loc["seg_iRiv"] = rnew[seg_r[sown]].astype(np.int32)
    loc["seg_length"] = np.asarray(mesh["seg_length"])[sown]
    loc["seg_Cwr"] = np.asarray(mesh["seg_Cwr"])[sown]
    # ---- lakes: kept only if wholly inside (cells and banks) ----
    ilake = np.asarray(mesh["ele_iLake"])
    if Nl and np.any(ilake[own] > 0):
        if np.any((ilake > 0) & ~owned_mask):
            raise NotImplementedError("a lake is cut by the partition")
        for k in ("lake_zmin", "lake_NumEleLake", "lake_bathy_ptr", "lake_bathy_yi", "lake_bathy_ai"):
            loc[k] = np.asarray(mesh[k])
        nl_loc = Nl
    else:
        loc["lake_zmin"] = np.zeros(0); loc["lake_NumEleLake"] = np.zeros(0, dtype=np.int32)
        loc["lake_bathy_ptr"] = np.zeros(1, dtype=np.int32); loc["lake_bathy_yi"] = np.zeros(0); loc["lake_bathy_ai"] = np.zeros(0)
        nl_loc = 0
    for k, v in (("Ne", nown), ("Nr", rown.size), ("Ns", sown.size), ("Nl", nl_loc),
                 ("close_boundary", int(np.asarray(mesh["close_boundary"]).reshape(-1)[0])),
                 ("lakeon", int(np.asarray(mesh["lakeon"]).reshape(-1)[0]) if nl_loc else 0)):
        loc[k] = np.array([v], dtype=np.int32)
    # ---- state vector, blocked ----
    if "y" in mesh:
        y = np.asarray(mesh["y"])
        loc["y"] = np.concatenate([y[own], y[Ne + own], y[2 * Ne + own], y[3 * Ne + rown],
                                   y[3 * Ne + Nr:3 * Ne + Nr + nl_loc]])
        # for tests: what the exchange must deliver, pair layout [Nhalo][2] = (Ysurf, Ygw)
        loc["halo_state_expected"] = np.stack([y[halo], y[2 * Ne + halo]], 1).ravel()
    loc["_own_ref"], loc["_riv_ref"] = own, rown
    return loc


def _reach_components(down):
    """component id of every reach of the river forest (reaches linked by riv_down > 0)"""
    down = np.asarray(down).astype(np.int64)
    n = down.size
    par = np.arange(n, dtype=np.int64)

    def find(a):
        while par[a] != a:
            par[a] = par[par[a]]
            a = par[a]
        return a

    for r in range(n):
        if down[r] > 0:
            a, b = find(r), find(down[r] - 1)
            if a != b:
                par[max(a, b)] = min(a, b)
    return np.array([find(r) for r in range(n)], dtype=np.int64)


def _hilbert_key(x, y, order=16):
    """Hilbert-curve index of points scaled into a 2^order grid (numpy, vectorised)"""
    x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    span = max(x.max() - x.min(), y.max() - y.min(), 1e-30)
    n = (1 << order) - 1
    xi = np.minimum(n, (x - x.min()) / span * n).astype(np.int64)
    yi = np.minimum(n, (y - y.min()) / span * n).astype(np.int64)
    d = np.zeros(xi.shape, dtype=np.int64)
    s = 1 << (order - 1)
    while s > 0:
        rx = (xi & s) > 0
        ry = (yi & s) > 0
        d += s * s * ((3 * rx.astype(np.int64)) ^ ry.astype(np.int64))
        flip = ~ry & rx
        xi = np.where(flip, n - xi, xi); yi = np.where(flip, n - yi, yi)
        swap = ~ry
        xi, yi = np.where(swap, yi, xi), np.where(swap, xi, yi)
        s >>= 1
    return d


def assign(mesh, nparts):
    """Owner rank of every cell for `nparts` partitions that `extract` accepts (Hilbert-range
    split on cell centroids, with the river and lake constraints of this version built in).  Cells that must stay
    together are merged into atoms first: the cells on the segments of one reach, a reach with its downstream reach
    (so a river tree and its banks form one atom), a lake with its cells, its bank cells and the reaches flowing
    into it, and a head-BC cell with its neighbours (a halo cell may not carry a head BC).  Atoms are then laid along
    the Hilbert curve of their centroids and cut into `nparts` runs of about Ne / nparts cells.
    Returns part_of_cell [Ne] (int32).  Balance is limited by the largest atom (one river tree = one partition)."""
    Ne, Nr, Ns, Nl = (int(np.asarray(mesh[k]).reshape(-1)[0]) for k in ("Ne", "Nr", "Ns", "Nl"))
    parent = np.arange(Ne, dtype=np.int64)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    def union(a, b):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)

    seg_e = np.asarray(mesh["seg_iEle"]).astype(np.int64) - 1
    seg_r = np.asarray(mesh["seg_iRiv"]).astype(np.int64) - 1
    comp = _reach_components(mesh["riv_down"]) if Nr else np.zeros(0, dtype=np.int64)
    rep = np.full(Nr, -1, dtype=np.int64)               # a representative cell of every river tree (by component id)
    for e, r in zip(seg_e, seg_r):
        c = comp[r]
        if rep[c] < 0:
            rep[c] = e
        else:
            union(e, rep[c])
    rep = rep[comp] if Nr else rep                      # ... seen from every reach of the tree (with or without segments)
    nabr = np.asarray(mesh["ele_nabr"]).reshape(3, Ne).astype(np.int64)
    ilake = np.asarray(mesh["ele_iLake"]).astype(np.int64)
    if Nl:
        lrep = np.full(Nl, -1, dtype=np.int64)
        for i in np.nonzero(ilake > 0)[0]:
            l = ilake[i] - 1
            if lrep[l] < 0:
                lrep[l] = i
            else:
                union(i, lrep[l])
        lnab = np.asarray(mesh["ele_lakenabr"]).reshape(3, Ne).astype(np.int64)
        for j in range(3):
            for i in np.nonzero(lnab[j] > 0)[0]:       # bank cells
                if lrep[lnab[j, i] - 1] >= 0:
                    union(i, lrep[lnab[j, i] - 1])
        tolake = np.asarray(mesh["riv_toLake"]).astype(np.int64) if "riv_toLake" in mesh else np.full(Nr, -1)
        for r in range(Nr):
            l = tolake[r]
            if 0 <= l < Nl and rep[r] >= 0 and lrep[l] >= 0:   # toLake is 0-based, -1 = none (ref_driver dump)
                union(rep[r], lrep[l])
    ibc = np.asarray(mesh["ele_iBC"])
    for i in np.nonzero(ibc > 0)[0]:
        for j in range(3):
            if nabr[j, i] > 0:
                union(i, nabr[j, i] - 1)
    root = np.array([find(i) for i in range(Ne)], dtype=np.int64)
    atoms, inv = np.unique(root, return_inverse=True)
    size = np.bincount(inv, minlength=atoms.size)
    x = np.asarray(mesh["ele_x"], dtype=np.float64); y = np.asarray(mesh["ele_y"], dtype=np.float64)
    ax = np.bincount(inv, weights=x, minlength=atoms.size) / size
    ay = np.bincount(inv, weights=y, minlength=atoms.size) / size
    order = np.argsort(_hilbert_key(ax, ay), kind="stable")
    part_of_atom = np.zeros(atoms.size, dtype=np.int32)
    target, acc, p = Ne / float(nparts), 0.0, 0
    for a in order:
        # move on when this partition is full, keeping at least one atom for every remaining partition
        if p < nparts - 1 and acc + 0.5 * size[a] > target * (p + 1):
            p += 1
        part_of_atom[a] = p
        acc += size[a]
    return part_of_atom[inv].astype(np.int32)


def cut_river_closure(mesh, part_of_cell, rank):
    """DESIGN of the next step (cut river trees) stated as data: for partition `rank` of an
    arbitrary cell partition, which cells and reaches it must hold so that owner-computes reproduces the single-domain
    result on what it owns, with no flux exchanged:
      own cells; halo cells (edge neighbours owned elsewhere: state Ysurf, Ygw);
      own reaches (a reach belongs to the partition owning most of its bank cells, ties to the lowest rank; a reach
      without segments follows its downstream reach, a tree without any segment goes to rank 0);
      replica cells (bank cells of own reaches that live elsewhere: full vertical parameters + Ysurf, Yunsat, Ygw; only
      their vertical role and segment fluxes are evaluated);
      halo reaches (reaches with a segment on an own cell, and the downstream / upstream reaches of own reaches that
      are owned elsewhere: statics + stage).
    Returns dict(own, halo, replica, riv_own, riv_halo, seg) of 0-based reference ids; `seg` = the segments between the
    held cells and the held reaches.  tests/test_partition_cut_design.py checks the closure with the CPU oracle; the
    CUDA path does not consume it yet (partition.extract still refuses cut reaches)."""
    Ne, Nr, Ns = (int(np.asarray(mesh[k]).reshape(-1)[0]) for k in ("Ne", "Nr", "Ns"))
    part = np.asarray(part_of_cell).astype(np.int64)
    nparts = int(part.max()) + 1
    seg_e = np.asarray(mesh["seg_iEle"]).astype(np.int64) - 1
    seg_r = np.asarray(mesh["seg_iRiv"]).astype(np.int64) - 1
    down = np.asarray(mesh["riv_down"]).astype(np.int64)
    # ---- owner of every reach ----
    votes = np.zeros((Nr, nparts), dtype=np.int64)
    np.add.at(votes, (seg_r, part[seg_e]), 1)
    has_seg = votes.sum(1) > 0
    riv_owner = np.where(has_seg, votes.argmax(1), -1)          # argmax: ties to the lowest rank
    for _ in range(Nr):                                         # segment-less reaches follow their downstream reach
        todo = np.nonzero((riv_owner < 0) & (down > 0))[0]
        if todo.size == 0:
            break
        new = riv_owner[down[todo] - 1]
        if np.all(new < 0):
            break
        riv_owner[todo] = np.where(new >= 0, new, -1)
    riv_owner[riv_owner < 0] = 0
    own = np.nonzero(part == rank)[0]
    nabr = np.asarray(mesh["ele_nabr"]).reshape(3, Ne).astype(np.int64)
    nb = nabr[:, own]
    nb0 = nb[nb > 0] - 1
    halo = np.unique(nb0[part[nb0] != rank])
    riv_own = np.nonzero(riv_owner == rank)[0]
    is_riv_own = np.zeros(Nr, dtype=bool); is_riv_own[riv_own] = True
    # replica cells: bank cells of own reaches that are not own cells
    bank = seg_e[is_riv_own[seg_r]]
    replica = np.unique(bank[part[bank] != rank])
    # halo reaches: reaches touching own cells, plus downstream / upstream neighbours of own reaches
    touch = np.unique(seg_r[part[seg_e] == rank])
    dn = down[riv_own]
    dn = dn[dn > 0] - 1
    ups = np.nonzero((down > 0) & is_riv_own[np.maximum(down - 1, 0)])[0]
    cand = np.unique(np.concatenate([touch, dn, ups]))
    riv_halo = cand[~is_riv_own[cand]]
    held_c = np.zeros(Ne, dtype=bool); held_c[own] = True; held_c[halo] = True; held_c[replica] = True
    held_r = np.zeros(Nr, dtype=bool); held_r[riv_own] = True; held_r[riv_halo] = True
    seg = np.nonzero(held_c[seg_e] & held_r[seg_r])[0] if Ns else np.zeros(0, dtype=np.int64)
    return dict(own=own, halo=halo, replica=replica, riv_own=riv_own, riv_halo=riv_halo, seg=seg, riv_owner=riv_owner)


def exchange_plan(own_gid, halo_gid, all_halo_gid):
    """Which of my cells each peer needs, and where what each peer sends lands in my halo arrays.
    all_halo_gid[q] = halo_gid of rank q (from an all_gather).  Both sides order a message by global id.
    returns send_ids[q] (local 0-based ids into my owned cells) and recv_pos[q] (0-based positions in my halo),
    where recv_pos[q] needs the peers' send lists: recv_from(q) = positions of (all_send_gid[q][me])."""
    order = np.argsort(own_gid, kind="stable")
    sorted_gid = own_gid[order]
    send_ids, send_gid = [], []
    for hq in all_halo_gid:
        hq = np.asarray(hq, dtype=np.int64)
        pos = np.searchsorted(sorted_gid, hq)
        pos = np.minimum(pos, max(sorted_gid.size - 1, 0))
        hit = (sorted_gid[pos] == hq) if sorted_gid.size else np.zeros(hq.size, dtype=bool)
        g = np.sort(hq[hit])
        send_gid.append(g)
        send_ids.append(order[np.searchsorted(sorted_gid, g)].astype(np.int64))
    return send_ids, send_gid


def recv_positions(halo_gid, gids_from_peer):
    """positions in my halo arrays of the cells a peer sends (it sends them ordered by global id)"""
    order = np.argsort(halo_gid, kind="stable")
    pos = np.searchsorted(halo_gid[order], np.asarray(gids_from_peer, dtype=np.int64))
    return order[pos].astype(np.int64)


class HaloExchange:
    """Per-RHS halo exchange over torch.distributed (NCCL on GPUs, gloo on CPU for tests).

    set-up: all_gather of the halo id lists -> send lists; all_to_all of the send id lists -> receive slots.
    each RHS: pack (Ysurf, Ygw) of my boundary cells (one gather kernel), one all_to_all_single, unpack into
    the halo buffers the kernels read (one index_copy each)."""

    def __init__(self, local, dist, device, cell_perm=None, pack_fn=None):
        import torch
        self.torch, self.dist, self.device = torch, dist, device
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        own_gid, halo_gid = np.asarray(local["own_gid"]), np.asarray(local["halo_gid"])
        self.Ne, self.Nhalo = own_gid.size, halo_gid.size
        if self.world > 1:
            all_halo = [None] * self.world
            dist.all_gather_object(all_halo, halo_gid)
        else:
            all_halo = [halo_gid]
        send_ids, send_gid = exchange_plan(own_gid, halo_gid, all_halo)
        send_ids[self.rank] = np.zeros(0, dtype=np.int64); send_gid[self.rank] = np.zeros(0, dtype=np.int64)
        if self.world > 1:
            all_send = [None] * self.world
            dist.all_gather_object(all_send, send_gid)
            recv_gid = [all_send[q][self.rank] if q != self.rank else np.zeros(0, dtype=np.int64) for q in range(self.world)]
        else:
            recv_gid = [np.zeros(0, dtype=np.int64)]
        self.send_counts = [int(s.size) for s in send_ids]
        self.recv_counts = [int(g.size) for g in recv_gid]
        assert sum(self.recv_counts) == self.Nhalo, (sum(self.recv_counts), self.Nhalo)
        ids = np.concatenate(send_ids) if send_ids else np.zeros(0, dtype=np.int64)
        self.send_cells_ref = ids.astype(np.int32)  # reference-local ids, concatenated by peer rank
        if cell_perm is not None:  # reference-local id -> device-order id of the context
            inv = np.empty(self.Ne, dtype=np.int64); inv[np.asarray(cell_perm)] = np.arange(self.Ne)
            ids = inv[ids]
        self.send_idx = torch.from_numpy(ids.astype(np.int32)).to(device)
        pos = np.concatenate([recv_positions(halo_gid, g) for g in recv_gid]) if self.Nhalo else np.zeros(0, dtype=np.int64)
        self.recv_pos = torch.from_numpy(pos).to(device)
        ns, nr = int(self.send_idx.numel()), self.Nhalo
        # pair layout: a cell travels as (Ysurf, Ygw); messages ordered by (owner, global id) land in place
        self.sbuf = torch.zeros(2 * max(ns, 1), dtype=torch.float64, device=device)
        self.rbuf = torch.zeros(2 * max(nr, 1), dtype=torch.float64, device=device)
        self.in_place = bool(nr == 0 or np.array_equal(pos, np.arange(nr)))
        self.side, self._work = None, None
        self.halo_state = self.rbuf if self.in_place else torch.zeros(2 * max(nr, 1), dtype=torch.float64, device=device)
        self.ns, self.nr = ns, nr
        self.pack_fn = pack_fn
        self.bytes_per_exchange = 16 * ns

    def native_plan(self):
        """(peer ranks, send counts, recv counts, send cells) for shud_b200_exchange_plan: the same exchange driven by
        the C library over its own NCCL communicator.  Needs the receives to land in halo order (they do when halo
        cells are numbered by (owner rank, global id), which partition.extract and synth stripes guarantee)."""
        if not self.in_place:
            raise NotImplementedError("halo cells are not ordered by (owner, global id)")
        peers = [q for q in range(self.world) if self.send_counts[q] or self.recv_counts[q]]
        return (np.asarray(peers, dtype=np.int32), np.asarray([self.send_counts[q] for q in peers], dtype=np.int32),
                np.asarray([self.recv_counts[q] for q in peers], dtype=np.int32), self.send_cells_ref)

    def start(self, y):
        """pack on the current stream and post the exchange on the side stream `self.side`: the caller overlaps it
        with the interior part of the RHS (ShudRHS.f_interior_dev), then calls finish() and hands the side stream to
        ShudRHS.f_boundary_dev, which runs the halo-dependent tiles there"""
        import torch
        ns, nr = self.ns, self.nr
        if self.pack_fn is not None:
            if ns:
                self.pack_fn(y, self.send_idx, self.sbuf)
        elif ns:
            idx = self.send_idx.long()
            self.sbuf[0:2 * ns:2] = y[idx]
            self.sbuf[1:2 * ns:2] = y[2 * self.Ne + idx]
        self._work = None
        if self.side is None and y.is_cuda:
            self.side = torch.cuda.Stream(device=y.device, priority=-1)
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream(y.device))
        if self.world > 1:
            if self.side is not None:
                with torch.cuda.stream(self.side):
                    self._work = self._post()
            else:
                self._work = self._post()

    def _post(self):
        return self.dist.all_to_all_single(self.rbuf[:2 * self.nr], self.sbuf[:2 * self.ns], [2 * c for c in self.recv_counts],
                                           [2 * c for c in self.send_counts], async_op=True)

    def finish(self):
        """order the side stream (CPU tensors: the caller) behind the posted exchange; returns the side stream, on
        which the halo state buffer is then valid (None on CPU)"""
        import contextlib
        import torch
        ctx = torch.cuda.stream(self.side) if self.side is not None else contextlib.nullcontext()
        with ctx:
            if self._work is not None:
                self._work.wait()
                self._work = None
            if self.nr and not self.in_place:
                self.halo_state.view(-1, 2)[:self.nr].index_copy_(0, self.recv_pos, self.rbuf[:2 * self.nr].view(-1, 2))
        return self.side

    def exchange(self, y):
        """y: my state vector [3 Ne + Nr + Nl] (device order of the context when pack_fn is the CUDA pack).
        Returns the halo state buffer [Nhalo][2] the kernels read."""
        torch, dist = self.torch, self.dist
        ns, nr = self.ns, self.nr
        if self.pack_fn is not None:
            if ns:
                self.pack_fn(y, self.send_idx, self.sbuf)       # CUDA gather kernel of the C ABI
        elif ns:
            idx = self.send_idx.long()
            self.sbuf[0:2 * ns:2] = y[idx]
            self.sbuf[1:2 * ns:2] = y[2 * self.Ne + idx]
        if self.world > 1:
            dist.all_to_all_single(self.rbuf[:2 * nr], self.sbuf[:2 * ns], [2 * c for c in self.recv_counts],
                                   [2 * c for c in self.send_counts])
        if nr and not self.in_place:
            self.halo_state.view(-1, 2)[:nr].index_copy_(0, self.recv_pos, self.rbuf[:2 * nr].view(-1, 2))
        return self.halo_state


# ------------------------------------------------------------------------------------------------------------------
# Cut river trees: any cell partition that keeps a lake with its cells and
# banks, and a head-BC cell with its neighbours, in one piece.  Owner-computes, nothing but STATES crosses a cut:
#   halo cells    edge neighbours owned elsewhere                       (Ysurf, Ygw)            as before
#   ghost cells   bank cells of my reaches that live elsewhere          (Ysurf, Yunsat, Ygw)    ordinary local cells, no
#                 edges; their vertical role and segment fluxes are evaluated here, their ydot is 0
#   ghost reaches reaches on my cells' banks / up- / downstream of my reaches / flowing into my lakes, owned elsewhere
#                 (stage)                                                ordinary local reaches, ydot 0
# ------------------------------------------------------------------------------------------------------------------
def assign_cells(mesh, nparts, reach_weight=0.0, lake_cell_weight=0.25, return_work=False):
    """Hilbert-range owner of every cell with only the constraints the cut-river path still has: a lake stays with its
    cells and bank cells, a head-BC cell with its neighbours.  River trees are cut wherever the ranges fall, so the
    balance no longer depends on the size of the largest tree.
    reach_weight: work of one reach (its routing, its segments) in units of one cell's work; it is spread over the
    reach's bank cells - which decide who owns the reach - so that ranges rich in rivers get fewer cells.
    lake_cell_weight: work of a lake cell in the same units - the cell kernel does none of the soil, edge or segment
    work for it (fun_Ele_lakeVertical / lakeHorizon, MD_ElementFlux.cpp:2-23), so a lake, which has to stay whole,
    weighs little and the ranges balance WORK rather than cell counts (qhh's lake is 14 % of its cells).
    return_work: also return the work of every cell (for balance reports)."""
    Ne, Nl = (int(np.asarray(mesh[k]).reshape(-1)[0]) for k in ("Ne", "Nl"))
    wcell = np.ones(Ne, dtype=np.float64)
    if Nl:
        wcell[np.asarray(mesh["ele_iLake"]) > 0] = lake_cell_weight
    if reach_weight > 0.0 and int(np.asarray(mesh["Ns"]).reshape(-1)[0]) > 0:
        se = np.asarray(mesh["seg_iEle"]).astype(np.int64) - 1
        sr = np.asarray(mesh["seg_iRiv"]).astype(np.int64) - 1
        per_reach = np.bincount(sr, minlength=int(np.asarray(mesh["Nr"]).reshape(-1)[0])).astype(np.float64)
        np.add.at(wcell, se, reach_weight / per_reach[sr])
    parent = np.arange(Ne, dtype=np.int64)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    def union(a, b):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)

    nabr = np.asarray(mesh["ele_nabr"]).reshape(3, Ne).astype(np.int64)
    ilake = np.asarray(mesh["ele_iLake"]).astype(np.int64)
    if Nl:
        lrep = np.full(Nl, -1, dtype=np.int64)
        for i in np.nonzero(ilake > 0)[0]:
            l = ilake[i] - 1
            if lrep[l] < 0:
                lrep[l] = i
            else:
                union(i, lrep[l])
        lnab = np.asarray(mesh["ele_lakenabr"]).reshape(3, Ne).astype(np.int64)
        for j in range(3):
            for i in np.nonzero(lnab[j] > 0)[0]:
                if lrep[lnab[j, i] - 1] >= 0:
                    union(i, lrep[lnab[j, i] - 1])
    for i in np.nonzero(np.asarray(mesh["ele_iBC"]) > 0)[0]:
        for j in range(3):
            if nabr[j, i] > 0:
                union(i, nabr[j, i] - 1)
    root = np.array([find(i) for i in range(Ne)], dtype=np.int64) if (Nl or np.any(np.asarray(mesh["ele_iBC"]) > 0)) \
        else np.arange(Ne, dtype=np.int64)
    atoms, inv = np.unique(root, return_inverse=True)
    size = np.bincount(inv, minlength=atoms.size)
    work = np.bincount(inv, weights=wcell, minlength=atoms.size)
    x = np.asarray(mesh["ele_x"], dtype=np.float64); y = np.asarray(mesh["ele_y"], dtype=np.float64)
    ax = np.bincount(inv, weights=x, minlength=atoms.size) / size
    ay = np.bincount(inv, weights=y, minlength=atoms.size) / size
    order = np.argsort(_hilbert_key(ax, ay), kind="stable")
    csum = np.cumsum(work[order]) - 0.5 * work[order]
    part_of_atom = np.empty(atoms.size, dtype=np.int32)
    part_of_atom[order] = np.minimum((csum * nparts / float(work.sum())).astype(np.int64), nparts - 1)
    part = part_of_atom[inv].astype(np.int32)
    return (part, wcell) if return_work else part


def _closure_with_lakes(mesh, part, rank):
    cl = cut_river_closure(mesh, part, rank)
    Nl = int(np.asarray(mesh["Nl"]).reshape(-1)[0])
    if Nl and "riv_toLake" in mesh:
        ilake = np.asarray(mesh["ele_iLake"]).astype(np.int64)
        lake_owner = np.full(Nl, -1, dtype=np.int64)
        for l in range(Nl):
            cells = np.nonzero(ilake == l + 1)[0]
            if cells.size:
                lake_owner[l] = part[cells[0]]
        tl = np.asarray(mesh["riv_toLake"]).astype(np.int64)
        into_mine = np.nonzero((tl >= 0) & (tl < Nl) & (lake_owner[np.clip(tl, 0, Nl - 1)] == rank))[0]
        extra = into_mine[cl["riv_owner"][into_mine] != rank]
        cl["riv_halo"] = np.unique(np.concatenate([cl["riv_halo"], extra]))
        cl["lake_owner"] = lake_owner
    return cl


def extract_cut(mesh, part_of_cell, rank, closures=None):
    """Local mesh of partition `rank` of the cell partition `part_of_cell` with river trees cut (see above), and the
    exchange lists.  Returns (loc, plan):
      loc   snapshot-named local mesh: cells = own + ghost cells (the last loc['n_ghost_cells']), halo_* arrays as
            `extract` makes them, reaches = own + ghost reaches (the last loc['n_ghost_reaches']), segments between held
            cells and held reaches with an own cell or an own reach; y in the local blocked layout (ghost entries hold
            the owners' values at extraction time; the kernels never read them - the exchange delivers them)
      plan  dict(peers, send_counts [npeers][3], recv_counts [npeers][3], send_items): per neighbour and kind (0 halo
            pairs, 1 ghost-cell triples, 2 ghost-reach stages) how many doubles travel, and the flat indices into MY
            local blocked vector (reference-local order) of what I send, grouped by (peer, kind), each group in the
            receiver's order (by global id)."""
    Ne, Nr, Ns, Nl = (int(np.asarray(mesh[k]).reshape(-1)[0]) for k in ("Ne", "Nr", "Ns", "Nl"))
    part = np.asarray(part_of_cell).astype(np.int64)
    nparts = int(part.max()) + 1
    if closures is None:
        closures = [_closure_with_lakes(mesh, part, p) for p in range(nparts)]
    cl = closures[rank]
    riv_owner = cl["riv_owner"]
    own = cl["own"]
    rep = cl["replica"]
    rep = rep[np.lexsort((rep, part[rep]))]                  # ghost cells by (owner, global id)
    halo = cl["halo"]
    halo = halo[np.lexsort((halo, part[halo]))]              # halo cells by (owner, global id)
    rh = cl["riv_halo"]
    rh = rh[np.lexsort((rh, riv_owner[rh]))]                 # ghost reaches by (owner, global id)
    cells = np.concatenate([own, rep])
    rivs = np.concatenate([cl["riv_own"], rh])
    nown, nloc, nrown = own.size, cells.size, cl["riv_own"].size
    ibc, ilake = np.asarray(mesh["ele_iBC"]), np.asarray(mesh["ele_iLake"])
    if np.any(ibc[halo] > 0) or np.any(ilake[halo] > 0) or np.any(ilake[rep] > 0):
        raise NotImplementedError("halo / ghost cells with a head BC or inside a lake (use assign_cells)")
    cnew = np.zeros(Ne, dtype=np.int64); cnew[cells] = np.arange(1, nloc + 1)
    hnew = np.zeros(Ne, dtype=np.int64); hnew[halo] = nloc + np.arange(1, halo.size + 1)
    rnew = np.zeros(Nr, dtype=np.int64); rnew[rivs] = np.arange(1, rivs.size + 1)
    nabr = np.asarray(mesh["ele_nabr"]).reshape(3, Ne).astype(np.int64)
    lnab = np.asarray(mesh["ele_lakenabr"]).reshape(3, Ne).astype(np.int64)
    nb = nabr[:, own]
    nb0 = np.maximum(nb - 1, 0)
    is_own_nb = (nb > 0) & (part[nb0] == rank)
    nab_loc = np.where(nb > 0, np.where(is_own_nb, cnew[nb0], hnew[nb0]), 0)
    if np.any(lnab[:, own][(nb > 0) & ~is_own_nb] > 0):
        raise NotImplementedError("a lake bank is cut by the partition (use assign_cells)")
    loc = {}
    for k, v in mesh.items():
        v = np.asarray(v)
        if k in CELL_SKIP:
            continue
        if k in EDGE_KEYS:
            loc[k] = np.ascontiguousarray(v.reshape(3, Ne)[:, cells]).ravel()
        elif (k.startswith("ele_") or k in CELL_DYN) and v.ndim == 1 and v.shape[0] == Ne:
            loc[k] = v[cells]
        elif k.startswith("riv_") and v.ndim == 1 and v.shape[0] == Nr:
            loc[k] = v[rivs]
    loc["ele_nabr"] = np.concatenate([nab_loc, np.zeros((3, rep.size), dtype=np.int64)], axis=1).astype(np.int32).ravel()
    loc["ele_lakenabr"] = np.concatenate([lnab[:, own], np.zeros((3, rep.size), dtype=np.int64)], axis=1).astype(np.int32).ravel()
    for k in HALO_KEYS:
        loc["halo_" + k] = np.asarray(mesh["ele_" + k])[halo]
    # lakes: whole and local, or absent
    lake_keep = np.zeros(Nl, dtype=bool)
    if Nl:
        lake_keep = cl["lake_owner"] == rank if "lake_owner" in cl else np.array([np.any(ilake[own] == l + 1) for l in range(Nl)])
        if np.any((ilake > 0) & lake_keep[np.clip(ilake - 1, 0, Nl - 1)] & (part != rank)):
            raise NotImplementedError("a lake is cut by the partition (use assign_cells)")
    lnew = np.zeros(Nl, dtype=np.int64); lnew[lake_keep] = np.arange(1, int(lake_keep.sum()) + 1)
    nl_loc = int(lake_keep.sum())
    if Nl:
        il = loc["ele_iLake"].astype(np.int64)
        loc["ele_iLake"] = np.where(il > 0, lnew[np.clip(il - 1, 0, Nl - 1)], 0).astype(np.int32)
        ln = loc["ele_lakenabr"].astype(np.int64)
        loc["ele_lakenabr"] = np.where(ln > 0, lnew[np.clip(ln - 1, 0, Nl - 1)], 0).astype(np.int32)
        ptr = np.asarray(mesh["lake_bathy_ptr"]).astype(np.int64)
        keep = np.nonzero(lake_keep)[0]
        loc["lake_zmin"] = np.asarray(mesh["lake_zmin"])[keep]
        loc["lake_NumEleLake"] = np.asarray(mesh["lake_NumEleLake"])[keep]
        yi, ai, p2 = [], [], [0]
        for l in keep:
            yi.append(np.asarray(mesh["lake_bathy_yi"])[ptr[l]:ptr[l + 1]]); ai.append(np.asarray(mesh["lake_bathy_ai"])[ptr[l]:ptr[l + 1]])
            p2.append(p2[-1] + int(ptr[l + 1] - ptr[l]))
        loc["lake_bathy_yi"] = np.concatenate(yi) if yi else np.zeros(0)
        loc["lake_bathy_ai"] = np.concatenate(ai) if ai else np.zeros(0)
        loc["lake_bathy_ptr"] = np.asarray(p2, dtype=np.int32)
        if "riv_toLake" in mesh:
            tl = np.asarray(mesh["riv_toLake"]).astype(np.int64)[rivs]
            ok = (tl >= 0) & (tl < Nl)
            # a lake held elsewhere keeps "flows into a lake" (the routing formula) under the index one past the local lakes
            loc["riv_toLake"] = np.where(ok, np.where(lake_keep[np.clip(tl, 0, Nl - 1)], lnew[np.clip(tl, 0, Nl - 1)] - 1, nl_loc), tl).astype(np.int32)
    else:
        loc["lake_zmin"] = np.zeros(0); loc["lake_NumEleLake"] = np.zeros(0, dtype=np.int32)
        loc["lake_bathy_ptr"] = np.zeros(1, dtype=np.int32); loc["lake_bathy_yi"] = np.zeros(0); loc["lake_bathy_ai"] = np.zeros(0)
    down = np.asarray(mesh["riv_down"]).astype(np.int64)[rivs]
    dn_new = np.where(down > 0, rnew[np.maximum(down - 1, 0)], down)
    if np.any((down[:nrown] > 0) & (dn_new[:nrown] == 0)):
        raise AssertionError("closure: the downstream reach of an own reach is not held")
    loc["riv_down"] = np.where((down > 0) & (dn_new == 0), -3, dn_new).astype(np.int32)   # ghost reach, downstream not held
    seg_e = np.asarray(mesh["seg_iEle"]).astype(np.int64) - 1
    seg_r = np.asarray(mesh["seg_iRiv"]).astype(np.int64) - 1
    is_own_c = part[seg_e] == rank
    is_own_r = riv_owner[seg_r] == rank
    sg = np.nonzero(is_own_c | is_own_r)[0] if Ns else np.zeros(0, dtype=np.int64)
    if sg.size and (np.any(cnew[seg_e[sg]] == 0) or np.any(rnew[seg_r[sg]] == 0)):
        raise AssertionError("closure: a needed segment touches a cell or reach that is not held")
    loc["seg_iEle"] = cnew[seg_e[sg]].astype(np.int32)
    loc["seg_iRiv"] = rnew[seg_r[sg]].astype(np.int32)
    loc["seg_length"] = np.asarray(mesh["seg_length"])[sg]
    loc["seg_Cwr"] = np.asarray(mesh["seg_Cwr"])[sg]
    for k, v in (("Ne", nloc), ("Nr", rivs.size), ("Ns", sg.size), ("Nl", nl_loc),
                 ("close_boundary", int(np.asarray(mesh["close_boundary"]).reshape(-1)[0])),
                 ("lakeon", int(np.asarray(mesh["lakeon"]).reshape(-1)[0]) if (nl_loc or ("riv_toLake" in loc and np.any(loc["riv_toLake"] >= 0))) else 0),
                 ("n_ghost_cells", rep.size), ("n_ghost_reaches", rh.size)):
        loc[k] = np.array([v], dtype=np.int32)
    lakes_kept = np.nonzero(lake_keep)[0]
    if "y" in mesh:
        y = np.asarray(mesh["y"])
        loc["y"] = np.concatenate([y[cells], y[Ne + cells], y[2 * Ne + cells], y[3 * Ne + rivs], y[3 * Ne + Nr + lakes_kept]])
        loc["halo_state_expected"] = np.stack([y[halo], y[2 * Ne + halo]], 1).ravel()
    loc["_own_ref"], loc["_riv_ref"], loc["_lake_ref"] = own, cl["riv_own"], lakes_kept
    loc["own_gid"], loc["halo_gid"] = own, halo
    # ---- exchange lists ----
    peers, send_counts, recv_counts, items = [], [], [], []
    for q in range(nparts):
        if q == rank:
            continue
        cq = closures[q]
        # what q needs from me, in q's order (by global id within my ownership)
        hq = np.sort(cq["halo"][part[cq["halo"]] == rank])
        gq = np.sort(cq["replica"][part[cq["replica"]] == rank])
        rq = np.sort(cq["riv_halo"][riv_owner[cq["riv_halo"]] == rank])
        # what I receive from q
        rc = [2 * int(np.sum(part[halo] == q)), 3 * int(np.sum(part[rep] == q)), int(np.sum(riv_owner[rh] == q))]
        sc = [2 * hq.size, 3 * gq.size, rq.size]
        if sum(rc) == 0 and sum(sc) == 0:
            continue
        hi = cnew[hq] - 1; gi = cnew[gq] - 1; ri = rnew[rq] - 1
        it = [np.stack([hi, 2 * nloc + hi], 1).ravel(), np.stack([gi, nloc + gi, 2 * nloc + gi], 1).ravel(), 3 * nloc + ri]
        peers.append(q); send_counts.append(sc); recv_counts.append(rc); items.append(np.concatenate(it))
    plan = dict(peers=np.asarray(peers, dtype=np.int32), send_counts=np.asarray(send_counts, dtype=np.int32).reshape(-1, 3),
                recv_counts=np.asarray(recv_counts, dtype=np.int32).reshape(-1, 3),
                send_items=(np.concatenate(items) if items else np.zeros(0)).astype(np.int32))
    return loc, plan
