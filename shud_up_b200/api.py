"""Host-side mirror of the reference's interface for the hot path, on top of the C ABI
(include/shud_b200.h -> shud_up_b200/libshud_b200.so).

    rhs = ShudRHS(mesh_dict)            # after Model_Data::initialize()  (src/Model/shud.cpp:51)
    rhs.set_forcing(forcing_dict)       # after updateforcing()+ET()      (src/Model/shud.cpp:106-109)
    rhs.f(t, y, ydot)                   # int f(t, N_Vector y, N_Vector ydot, void*)  (src/Model/f.hpp:12)

The CUDA library is the only implementation: if it cannot be loaded, or there is no CUDA
device, construction raises - nothing here computes on the CPU.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SHUD_B200_LIB", os.path.join(_HERE, "libshud_b200.so"))  # env override: A/B builds
_lib = None
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)


class ShudError(RuntimeError):
    pass


# myexit() codes of the reference (src/Equations/functions.cpp:10-36, src/Model/Macros.hpp:77-82)
ERR_TEXT = {10: "NAN/INF VALUE (ERRNAN)", 13: "Data validation (ERRDATAIN)", 1: "River Routing Boundary Condition Type Is Wrong",
            -1: "no CUDA device", -2: "bad argument", -3: "CUDA runtime error"}


def lib():
    """Load libshud_b200.so (built in-tree by shud_up_b200/build.py).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ShudError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sig = {
        "shud_b200_create": (C.c_int, [C.POINTER(abi.ShudMesh), C.c_int, C.POINTER(vp)]),
        "shud_b200_create_partition": (C.c_int, [C.POINTER(abi.ShudMesh), C.POINTER(abi.ShudHalo), C.c_int, C.POINTER(vp)]),
        "shud_b200_set_halo_state_dev": (C.c_int, [vp, vp]),
        "shud_b200_pack_halo_dev": (C.c_int, [vp, vp, vp, C.c_int32, vp]),
        "shud_b200_destroy": (None, [vp]),
        "shud_b200_ny": (C.c_int64, [vp]),
        "shud_b200_stream": (vp, [vp]),
        "shud_b200_set_forcing": (C.c_int, [vp, C.POINTER(abi.ShudForcing)]),
        "shud_b200_prime": (C.c_int, [vp, _PD]),
        "shud_b200_set_carried": (C.c_int, [vp, _PD]),
        "shud_b200_get_carried": (C.c_int, [vp, _PD, _PD]),
        "shud_b200_to_device_order": (C.c_int, [vp, vp, vp]),
        "shud_b200_from_device_order": (C.c_int, [vp, vp, vp]),
        "shud_b200_summary_dev": (C.c_int, [vp, vp, _PD]),
        "shud_b200_exchange_plan_items": (C.c_int, [vp, C.c_int, _PI, _PI, _PI, _PI]),
        "shud_b200_p2p_export": (C.c_int, [vp, C.c_int, vp]),
        "shud_b200_p2p_connect": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "shud_b200_allreduce": (C.c_int, [vp, _PD, C.c_int, C.c_int]),
        "shud_b200_perm": (C.c_int, [vp, _PI, _PI]),
        "shud_b200_rhs_dev": (C.c_int, [vp, C.c_double, vp, vp]),
        "shud_b200_rhs_dq_dev": (C.c_int, [vp, C.c_double, C.c_double, vp, vp, vp, vp, vp, vp, vp]),
        "shud_b200_rhs": (C.c_int, [vp, C.c_double, vp, vp]),
        "shud_b200_rhs_stage_dev": (C.c_int, [vp, C.c_int, vp, vp]),
        "shud_b200_rhs_interior_dev": (C.c_int, [vp, C.c_double, vp, vp]),
        "shud_b200_rhs_boundary_dev": (C.c_int, [vp, C.c_double, vp, vp, vp]),
        "shud_b200_tile_counts": (C.c_int, [vp, _PI, _PI]),
        "shud_b200_mesh_save": (C.c_int, [C.c_char_p, C.POINTER(abi.ShudMesh)]),
        "shud_b200_mesh_load": (C.c_int, [C.c_char_p, C.POINTER(abi.ShudMesh), C.POINTER(C.c_void_p)]),
        "shud_b200_mesh_free": (None, [C.c_void_p]),
        "shud_b200_format_ic": (C.c_int, [C.c_char_p, C.c_double, C.c_int32, C.c_int32, C.c_int32, _PD, _PD, _PD]),
        "shud_b200_write_ic": (C.c_int, [vp, C.c_char_p, C.c_double, vp]),
        "shud_b200_read_ic": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, _PD, _PD, _PD, _PD]),
        "shud_b200_land_create": (C.c_int, [vp, C.POINTER(abi.ShudLand)]),
        "shud_b200_land_set_state": (C.c_int, [vp, _PD, _PD]),
        "shud_b200_land_step": (C.c_int, [vp, C.POINTER(abi.ShudLandStep)]),
        "shud_b200_land_get": (C.c_int, [vp, C.POINTER(abi.ShudLandOut)]),
        "shud_b200_comm_unique_id": (C.c_int, [C.c_char_p, vp]),
        "shud_b200_comm_init": (C.c_int, [vp, C.c_char_p, vp, C.c_int, C.c_int]),
        "shud_b200_exchange_plan": (C.c_int, [vp, C.c_int, _PI, _PI, _PI, _PI]),
        "shud_b200_rhs_exchange_dev": (C.c_int, [vp, C.c_double, vp, vp]),
        "shud_b200_rhs_diag_dev": (C.c_int, [vp, C.c_double, vp, vp]),
        "shud_b200_get_diag": (C.c_int, [vp, C.POINTER(abi.ShudDiag)]),
        "shud_b200_output_accumulate": (C.c_int, [vp]),
        "shud_b200_output_flush": (C.c_int, [vp, C.c_double, C.POINTER(abi.ShudDiag), _PI]),
        "shud_b200_check": (C.c_int, [vp, _PI]),
        "shud_b200_launches_per_rhs": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        if "SHUD_B200_LIB" in os.environ and not hasattr(L, name):
            continue  # an A/B build of an older source may lack newer entry points (developer knob only)
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _chk(rc, what):
    if rc != 0:
        raise ShudError(f"{what} failed: code {rc} ({ERR_TEXT.get(rc, '?')})")


def _ptr(a):
    """device/host pointer of a torch tensor or numpy array (float64, contiguous)."""
    if hasattr(a, "data_ptr"):
        assert a.dtype.is_floating_point and a.element_size() == 8 and a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return C.c_void_p(a.ctypes.data)


class ShudRHS:
    """One GPU's copy of the model: static SoA mirror + forcing-step arrays + carried state."""

    def __init__(self, mesh, device=0):
        """mesh: dict of the static SoA arrays; if it carries halo_* arrays it is one partition of a
        larger mesh (shud_up_b200/partition.py) and nabr may name halo cells Ne+1..Ne+Nhalo."""
        L = lib()
        if isinstance(mesh, LoadedMesh):  # a binary container read by shud_b200_mesh_load (no halo: whole domain)
            self._mesh_struct, self._keep, mesh = mesh.mesh, [mesh], {}
        else:
            self._mesh_struct, self._keep = abi.make_mesh(mesh)
        self.Ne, self.Nr, self.Ns, self.Nl = (self._mesh_struct.Ne, self._mesh_struct.Nr, self._mesh_struct.Ns,
                                              self._mesh_struct.Nl)
        h = C.c_void_p()
        self.Nhalo = 0
        ghosts = any(int(np.asarray(mesh[k]).reshape(-1)[0]) > 0 for k in ("n_ghost_cells", "n_ghost_reaches") if k in mesh)
        if "halo_z_surf" in mesh and (len(mesh["halo_z_surf"]) > 0 or ghosts):
            hs, keep = abi.make_halo(mesh)
            self._keep += keep
            self.Nhalo = hs.Nhalo
            rc = L.shud_b200_create_partition(C.byref(self._mesh_struct), C.byref(hs), int(device), C.byref(h))
        else:
            rc = L.shud_b200_create(C.byref(self._mesh_struct), int(device), C.byref(h))
        _chk(rc, "shud_b200_create")
        self._h = h
        self.device = int(device)
        self.NY = int(L.shud_b200_ny(h))
        self.launches_per_rhs = int(L.shud_b200_launches_per_rhs(h))

    def close(self):
        if getattr(self, "_h", None):
            lib().shud_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ----
    @property
    def stream_ptr(self):
        return lib().shud_b200_stream(self._h)

    def torch_stream(self):
        import torch
        return torch.cuda.ExternalStream(self.stream_ptr, device=f"cuda:{self.device}")

    def perm(self):
        cp = np.empty(self.Ne, dtype=np.int32)
        rp = np.empty(max(self.Nr, 1), dtype=np.int32)
        _chk(lib().shud_b200_perm(self._h, cp.ctypes.data_as(_PI), rp.ctypes.data_as(_PI)), "perm")
        return cp, rp[:self.Nr]

    def to_device_order(self, ref_dev, out_dev):
        _chk(lib().shud_b200_to_device_order(self._h, _ptr(ref_dev), _ptr(out_dev)), "to_device_order")

    def from_device_order(self, dev_dev, out_ref):
        _chk(lib().shud_b200_from_device_order(self._h, _ptr(dev_dev), _ptr(out_ref)), "from_device_order")

    def summary(self, y_dev):
        """Model_Data::summary (MD_update.cpp:190-216): the device vector as a host array in reference order, BC heads
        and BC stages in place of the solver's frozen rows"""
        out = np.empty(self.NY)
        _chk(lib().shud_b200_summary_dev(self._h, _ptr(y_dev), out.ctypes.data_as(_PD)), "summary")
        return out

    # ---- halo exchange plumbing (multi-GPU) ----
    def set_halo_state(self, halo_state_dev):
        """register the device buffer [Nhalo][2] = (Ysurf, Ygw) the exchange fills before each RHS"""
        self._halo_buf = halo_state_dev
        _chk(lib().shud_b200_set_halo_state_dev(self._h, _ptr(halo_state_dev)), "set_halo_state")

    def pack_halo(self, y_dev, idx_dev, out_dev):
        """out[2k] = Ysurf[idx[k]], out[2k+1] = Ygw[idx[k]]  (idx: int32 device-order cell ids, on the device)"""
        n = int(idx_dev.numel())
        _chk(lib().shud_b200_pack_halo_dev(self._h, _ptr(y_dev), C.c_void_p(idx_dev.data_ptr()), n, _ptr(out_dev)),
             "pack_halo")

    # ---- forcing step / carried state ----
    def set_forcing(self, forcing, qEleE_IC=None):
        f, keep = abi.make_forcing(forcing, qEleE_IC=qEleE_IC)
        _chk(lib().shud_b200_set_forcing(self._h, C.byref(f)), "shud_b200_set_forcing")

    def prime(self, y_host):
        y = np.ascontiguousarray(y_host, dtype=np.float64)
        _chk(lib().shud_b200_prime(self._h, y.ctypes.data_as(_PD)), "shud_b200_prime")

    def set_carried(self, u_satn):
        a = np.ascontiguousarray(u_satn, dtype=np.float64)
        _chk(lib().shud_b200_set_carried(self._h, a.ctypes.data_as(_PD)), "shud_b200_set_carried")

    def get_carried(self):
        s = np.empty(self.Ne)
        e = np.empty(self.Ne)
        _chk(lib().shud_b200_get_carried(self._h, s.ctypes.data_as(_PD), e.ctypes.data_as(_PD)), "get_carried")
        return s, e

    # ---- the RHS ----
    def f(self, t, y, ydot):
        """CVRhsFn on host vectors in reference order (numpy, or pinned torch CPU tensors).
        Raises ShudError with the reference's exit code where the reference would myexit()."""
        rc = lib().shud_b200_rhs(self._h, float(t), _ptr(y), _ptr(ydot))
        if rc != 0:
            raise ShudError(f"f(): reference would exit with code {rc} ({ERR_TEXT.get(rc, '?')})")
        return 0

    def f_dev(self, t, y_dev, ydot_dev, diag=False):
        """asynchronous RHS on device vectors in device order (torch cuda float64 tensors)."""
        fn = lib().shud_b200_rhs_diag_dev if diag else lib().shud_b200_rhs_dev
        _chk(fn(self._h, float(t), _ptr(y_dev), _ptr(ydot_dev)), "shud_b200_rhs_dev")

    def f_dq_dev(self, t, sigma, v_dev, ewt_dev, y0_dev, ytemp_dev, ydot_dev, ss_dev=None, vnorm_dev=None):
        """ytemp = sigma (v ./ ewt) + y0 formed by the pre-pass, ydot = f(t, ytemp): the difference-quotient evaluation
        of SPGMR's J v (single domain).  ss_dev: v is unnormalised, its squared 2-norm lies in ss_dev[0]; the
        normalised direction goes to vnorm_dev"""
        _chk(lib().shud_b200_rhs_dq_dev(self._h, float(t), float(sigma), _ptr(v_dev), _ptr(ewt_dev), _ptr(y0_dev),
                                        _ptr(ytemp_dev), _ptr(ydot_dev), _ptr(ss_dev) if ss_dev is not None else None,
                                        _ptr(vnorm_dev) if vnorm_dev is not None else None), "shud_b200_rhs_dq_dev")

    def f_interior_dev(self, t, y_dev, ydot_dev):
        """part of the RHS of a partition that needs no exchanged halo data (overlaps the halo exchange)"""
        _chk(lib().shud_b200_rhs_interior_dev(self._h, float(t), _ptr(y_dev), _ptr(ydot_dev)), "rhs_interior_dev")

    def f_boundary_dev(self, t, y_dev, ydot_dev, halo_stream=None):
        """the rest of the RHS, after the halo exchange has landed; `halo_stream` = the torch stream the exchange
        completes on (its halo-dependent tiles run there, beside the interior tiles), None = the context stream"""
        hs = C.c_void_p(halo_stream.cuda_stream) if halo_stream is not None else None
        _chk(lib().shud_b200_rhs_boundary_dev(self._h, float(t), _ptr(y_dev), _ptr(ydot_dev), hs), "rhs_boundary_dev")

    # ---- land-surface step on the device (updateforcing + ET of the reference) ----
    def land_create(self, land):
        """land: abi.ShudLand (abi.make_land builds it from a snapshot)"""
        _chk(lib().shud_b200_land_create(self._h, C.byref(land)), "land_create")

    def land_set_state(self, yEleSnow, yEleIS):
        a = np.ascontiguousarray(yEleSnow, dtype=np.float64); b = np.ascontiguousarray(yEleIS, dtype=np.float64)
        _chk(lib().shud_b200_land_set_state(self._h, a.ctypes.data_as(_PD), b.ctypes.data_as(_PD)), "land_set_state")

    def land_step(self, step):
        """step: abi.ShudLandStep; writes the RHS's forcing arrays on the device (no set_forcing needed)"""
        _chk(lib().shud_b200_land_step(self._h, C.byref(step)), "land_step")

    def land_get(self):
        o, arrs = abi.make_land_out(self.Ne)
        _chk(lib().shud_b200_land_get(self._h, C.byref(o)), "land_get")
        return arrs

    # ---- halo exchange driven by the library over its own NCCL communicator ----
    @staticmethod
    def nccl_library():
        """libnccl.so.2 of the running process (the copy torch has loaded, so no second NCCL enters the process)"""
        try:
            import nvidia.nccl
            for d in nvidia.nccl.__path__:
                p = os.path.join(d, "lib", "libnccl.so.2")
                if os.path.exists(p):
                    return p
        except ImportError:
            pass
        return "libnccl.so.2"

    def comm_init(self, dist, device, nccl_lib=None):
        """collective over the ranks of `dist` (torch.distributed: used once, to hand rank 0's id to the others)"""
        import torch
        path = (nccl_lib or self.nccl_library()).encode()
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            _chk(lib().shud_b200_comm_unique_id(path, buf), "comm_unique_id")
        t = torch.tensor(list(buf), dtype=torch.uint8, device=device)
        dist.broadcast(t, 0)
        raw = (C.c_ubyte * 128)(*[int(v) for v in t.cpu().tolist()])
        _chk(lib().shud_b200_comm_init(self._h, path, raw, rank, world), "comm_init")

    def exchange_plan(self, peers, send_counts, recv_counts, send_cells):
        self._last_plan = (peers, send_counts, recv_counts, send_cells)
        a = [np.ascontiguousarray(v, dtype=np.int32) for v in (peers, send_counts, recv_counts, send_cells)]
        ptr = [v.ctypes.data_as(_PI) for v in a]
        _chk(lib().shud_b200_exchange_plan(self._h, int(a[0].size), *ptr), "exchange_plan")

    def exchange_plan_items(self, plan):
        """plan: partition.extract_cut's dict (peers, send_counts [n][3], recv_counts [n][3], send_items): the general
        exchange - halo pairs, ghost-cell triples, ghost-reach stages - of the peer-to-peer path"""
        a = [np.ascontiguousarray(plan[k], dtype=np.int32).ravel() for k in ("peers", "send_counts", "recv_counts", "send_items")]
        self._last_plan = None
        _chk(lib().shud_b200_exchange_plan_items(self._h, int(a[0].size), *[v.ctypes.data_as(_PI) for v in a]),
             "exchange_plan_items")

    P2P_BLOB = 512  # SHUD_P2P_BLOB_BYTES

    def p2p_export(self, rank):
        buf = (C.c_ubyte * self.P2P_BLOB)()
        _chk(lib().shud_b200_p2p_export(self._h, int(rank), buf), "p2p_export")
        return bytes(buf)

    def p2p_connect_blobs(self, rank, blobs):
        """blobs: list of every rank's p2p_export() bytes, rank order.  False: peer mapping unavailable (NCCL path stays)"""
        raw = b"".join(blobs)
        rc = lib().shud_b200_p2p_connect(self._h, int(rank), len(blobs), raw)
        if rc == -3:
            return False
        _chk(rc, "p2p_connect")
        return True

    def p2p_mailboxes(self):
        """(nranks, rank, [device pointers]) of the in-kernel allreduce mailboxes mapped by p2p_connect (nranks 0: none)"""
        nr, rk, bx = C.c_int(0), C.c_int(0), (C.c_void_p * 16)()
        fn = lib().shud_b200_p2p_mailboxes
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
        _chk(fn(self._h, C.byref(nr), C.byref(rk), bx), "p2p_mailboxes")
        return nr.value, rk.value, [bx[k] for k in range(nr.value)]

    def p2p_connect(self, dist, device):
        """collective: exchange the halo-buffer descriptors of all ranks, map the neighbours' buffers (CUDA IPC over
        NVLink), barrier.  After it f_exchange_dev moves the halo with peer stores + flags instead of NCCL."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        mine = torch.tensor(list(self.p2p_export(rank)), dtype=torch.uint8, device=device)
        allb = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine)
        ok = self.p2p_connect_blobs(rank, [bytes(t.cpu().tolist()) for t in allb])
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)   # all ranks or none
        if not int(flag.item()) and ok:
            self.exchange_plan(*self._last_plan)       # back to the NCCL buffers
        dist.barrier()
        return bool(int(flag.item()))

    def f_exchange_dev(self, t, y_dev, ydot_dev):
        """one f() of a partition: pack, NCCL sends/receives, interior part beside them, boundary part"""
        _chk(lib().shud_b200_rhs_exchange_dev(self._h, float(t), _ptr(y_dev), _ptr(ydot_dev)), "rhs_exchange_dev")

    def write_ic(self, path, t, y_dev):
        """checkpoint in the reference's <prj>.cfg.ic.update format, straight from the device vector"""
        _chk(lib().shud_b200_write_ic(self._h, str(path).encode(), float(t), _ptr(y_dev)), "write_ic")

    def tile_counts(self):
        """(interior, boundary) 128-cell tiles of this partition"""
        a, b = C.c_int(0), C.c_int(0)
        _chk(lib().shud_b200_tile_counts(self._h, C.byref(a), C.byref(b)), "tile_counts")
        return a.value, b.value

    def f_stage_dev(self, stage, y_dev, ydot_dev):
        _chk(lib().shud_b200_rhs_stage_dev(self._h, int(stage), _ptr(y_dev), _ptr(ydot_dev)), "rhs_stage_dev")

    def check(self):
        where = C.c_int32(0)
        code = lib().shud_b200_check(self._h, C.byref(where))
        return code, where.value

    def output_accumulate(self):
        """Print_Ctrl::PrintData's `buffer += value` for every flux array, on the device (after a diag RHS)"""
        _chk(lib().shud_b200_output_accumulate(self._h), "shud_b200_output_accumulate")

    def output_flush(self, tau):
        """interval means * tau of every flux array (host, reference order); resets the device buffers"""
        d, arrs = abi.make_diag(self.Ne, self.Nr, self.Ns, self.Nl)
        n = C.c_int32(0)
        _chk(lib().shud_b200_output_flush(self._h, float(tau), C.byref(d), C.byref(n)), "shud_b200_output_flush")
        return arrs, n.value

    def get_diag(self):
        d, arrs = abi.make_diag(self.Ne, self.Nr, self.Ns, self.Nl)
        _chk(lib().shud_b200_get_diag(self._h, C.byref(d)), "shud_b200_get_diag")
        return arrs


# ---- host-side ingest / checkpoint helpers (no device needed) ----
def mesh_save(path, snap):
    """write the shud_mesh SoA of a snapshot dict as one binary container"""
    m, keep = abi.make_mesh(snap)
    _chk(lib().shud_b200_mesh_save(str(path).encode(), C.byref(m)), "mesh_save")


class LoadedMesh:
    """a mesh container read back: .mesh is the abi.ShudMesh pointing into one block (freed by close())"""

    def __init__(self, path):
        self.mesh, self._block = abi.ShudMesh(), C.c_void_p()
        _chk(lib().shud_b200_mesh_load(str(path).encode(), C.byref(self.mesh), C.byref(self._block)), "mesh_load")

    def array(self, name, count, dtype=np.float64):
        p = getattr(self.mesh, name)
        if not p:
            return None
        return np.ctypeslib.as_array(p, shape=(count,)).copy()

    def close(self):
        if self._block:
            lib().shud_b200_mesh_free(self._block)
            self._block = C.c_void_p()


def format_ic(path, t, Ne, Nr, Nl, y, yEleIS=None, yEleSnow=None):
    y = np.ascontiguousarray(y, dtype=np.float64)
    a = None if yEleIS is None else np.ascontiguousarray(yEleIS, dtype=np.float64)
    b = None if yEleSnow is None else np.ascontiguousarray(yEleSnow, dtype=np.float64)
    _chk(lib().shud_b200_format_ic(str(path).encode(), float(t), int(Ne), int(Nr), int(Nl),
                                   a.ctypes.data_as(_PD) if a is not None else None,
                                   b.ctypes.data_as(_PD) if b is not None else None, y.ctypes.data_as(_PD)), "format_ic")


def read_ic(path, Ne, Nr, Nl):
    """-> (t, y, yEleIS, yEleSnow) from a <prj>.cfg.ic(.update) file"""
    y = np.empty(3 * Ne + Nr + Nl); a = np.empty(Ne); b = np.empty(Ne); t = C.c_double(0.0)
    _chk(lib().shud_b200_read_ic(str(path).encode(), int(Ne), int(Nr), int(Nl), C.byref(t), a.ctypes.data_as(_PD),
                                 b.ctypes.data_as(_PD), y.ctypes.data_as(_PD)), "read_ic")
    return t.value, y, a, b
