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
