"""One rank of tests/test_dist_nk_gpu.py: the library's CVODE-shaped integrator with the device-fused hooks on a
partition of ccw whose rivers are cut (ghost cells / reaches), as its own process on cuda:0.  Halo exchange: peer stores +
flags into the other process's block (CUDA IPC); every reduction of the distributed vector: allreduce inside the
reduction kernel through the mailboxes.  world = 1: the single domain, same integrator, for comparison.
Writes <dir>/nk<rank>.npz: owned entries of the end state with their global ids, integrator statistics."""
import ctypes as C
import os, sys, time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import torch
import oracle_lib
from shud_up_b200 import cvode as _cv, partition
from shud_up_b200.api import ShudRHS, lib as _lib

rank, world, d, t_end = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], float(sys.argv[4])
mode = sys.argv[5] if len(sys.argv) > 5 else "cut"  # cut: Hilbert ranges through the river network; trees: whole river trees per rank


def wait_for(path, timeout=180.0):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise SystemExit(f"rank {rank}: {path} did not appear")
        time.sleep(0.01)


def publish(name, data=b""):
    with open(os.path.join(d, name + ".tmp"), "wb") as f:
        f.write(data)
    os.rename(os.path.join(d, name + ".tmp"), os.path.join(d, name))


whole = dict(oracle_lib.load_case("ccw", "rand1"))  # a state away from equilibrium: the steps stay short for a while
Ne, Nr, Nl = int(whole["Ne"][0]), int(whole["Nr"][0]), int(whole["Nl"][0])
if world == 1:
    loc, plan = whole, None
    own_c, own_r, own_l = np.arange(Ne), np.arange(Nr), np.arange(Nl)
else:
    part = partition.assign_cells(whole, world) if mode == "cut" else partition.assign(whole, world)
    closures = [partition._closure_with_lakes(whole, part, p) for p in range(world)]
    loc, plan = partition.extract_cut(whole, part, rank, closures)
    own_c, own_r, own_l = loc["_own_ref"], loc["_riv_ref"], loc["_lake_ref"]
try:
    torch.zeros(1, device="cuda")
except Exception as e:  # a GPU in exclusive-process mode admits one context: the test is skipped, not failed
    print("NO_SECOND_CONTEXT", e)
    raise SystemExit(77)
rhs = ShudRHS(loc)
rhs.set_forcing(loc, qEleE_IC=loc["qEleE_IC_in"])
rhs.prime(loc["y"])
n_glob = rhs.NY
if world > 1:
    rhs.exchange_plan_items(plan)
    publish(f"blob{rank}", rhs.p2p_export(rank))
    blobs = []
    for r in range(world):
        wait_for(os.path.join(d, f"blob{r}"))
        blobs.append(open(os.path.join(d, f"blob{r}"), "rb").read())
    assert rhs.p2p_connect_blobs(rank, blobs)
    # the global length counts OWNED entries only: a ghost entry (ydot = 0, state from the exchange) is zero in every
    # correction, residual and Krylov vector, so it adds nothing to a sum - it must not add to N either
    publish(f"ny{rank}", str(3 * own_c.size + own_r.size + own_l.size).encode())
    n_glob = 0
    for r in range(world):
        wait_for(os.path.join(d, f"ny{r}"))
        n_glob += int(open(os.path.join(d, f"ny{r}")).read())
    assert rhs.p2p_mailboxes()[0] == world

L = _cv.bind(_lib())
L.N_VNew_ShudB200.restype = C.c_void_p
L.N_VNew_ShudB200.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
L.N_VCopyToDevice_ShudB200.argtypes = [C.c_void_p]
L.N_VSetDistributed_ShudB200.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
ws = C.c_void_p()
assert L.shud_nv_ws_create(0, C.c_void_p(rhs.stream_ptr), C.byref(ws)) == 0
yv = C.c_void_p(L.N_VNew_ShudB200(rhs.NY, ws, rhs._h, None))
np.ctypeslib.as_array(L.N_VGetArrayPointer(yv), shape=(rhs.NY,))[:] = loc["y"]
assert L.N_VCopyToDevice_ShudB200(yv) == 0
if world > 1:
    L.N_VSetDistributed_ShudB200(yv, n_glob, C.c_void_p(_cv.fn_address(L, "shud_b200_nv_allreduce")), rhs._h)
cvi = _cv.CVode(L, _cv.fn_address(L, "shud_b200_f_exchange" if world > 1 else "shud_b200_f"), rhs._h.value, 0.0, yv)
cvi.configure(rtol=1e-4, atol=1e-4, init_step=1e-3, max_step=0.5)
fz = _cv.Fused()
L.shud_b200_cv_fused_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(_cv.Fused)]
L.shud_b200_cv_fused_destroy.argtypes = [C.POINTER(_cv.Fused)]
assert L.shud_b200_cv_fused_create(rhs._h, ws, 5, C.byref(fz)) == 0
cvi.set_fused(fz)
if world > 1:  # everybody is mapped and set up before the first kernel touches a peer's block
    publish(f"ready{rank}")
    for r in range(world):
        wait_for(os.path.join(d, f"ready{r}"))
t0 = time.time()
cvi.solve(t_end, yv)
wall = time.time() - t0
assert rhs.check()[0] == 0
st = cvi.stats()
y = np.ctypeslib.as_array(L.N_VGetArrayPointer(yv), shape=(rhs.NY,)).copy()
nloc = rhs.Ne
gid = np.concatenate([b * Ne + own_c for b in range(3)] + [3 * Ne + own_r, 3 * Ne + Nr + own_l])
val = np.concatenate([y[b * nloc:b * nloc + own_c.size] for b in range(3)] +
                     [y[3 * nloc:3 * nloc + own_r.size], y[3 * nloc + rhs.Nr:3 * nloc + rhs.Nr + own_l.size]])
np.savez(os.path.join(d, f"nk{rank}.tmp.npz"), gid=gid, val=val, wall=wall, n_glob=n_glob, ny=rhs.NY,
         **{k: np.array([v]) for k, v in st.items()})
os.rename(os.path.join(d, f"nk{rank}.tmp.npz"), os.path.join(d, f"nk{rank}.npz"))
if world > 1:  # nobody unmaps while a peer may still be inside f() or a reduction
    for r in range(world):
        wait_for(os.path.join(d, f"nk{r}.npz"))
cvi.close()
L.shud_b200_cv_fused_destroy(C.byref(fz))
L.N_VDestroy(yv)
L.shud_nv_ws_destroy(ws)
print("rank", rank, "ok", st["nst"], wall)
