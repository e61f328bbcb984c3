"""One rank of tests/test_allreduce_gpu.py: a context on cuda:0, the ranks' peer-to-peer blocks exchanged through files
and mapped with CUDA IPC, a vector workspace whose reductions are combined by the reduction kernels themselves through
the mailboxes.  Writes the local and the global value of every reduction to <dir>/result<rank>.npz."""
import os, sys, time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import torch
import oracle_lib
from shud_up_b200.api import ShudRHS
from shud_up_b200.nvector import NVectorOps

rank, world, d = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]


def wait_for(path, timeout=120.0):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise SystemExit(f"rank {rank}: {path} did not appear")
        time.sleep(0.01)


try:
    torch.zeros(1, device="cuda")
except Exception as e:  # a GPU in exclusive-process mode admits one context: the test is skipped, not failed
    print("NO_SECOND_CONTEXT", e)
    raise SystemExit(77)
rhs = ShudRHS(oracle_lib.load_case("ccw", "ic"))
blob = rhs.p2p_export(rank)
with open(os.path.join(d, f"blob{rank}.tmp"), "wb") as f:
    f.write(blob)
os.rename(os.path.join(d, f"blob{rank}.tmp"), os.path.join(d, f"blob{rank}"))
blobs = []
for r in range(world):
    wait_for(os.path.join(d, f"blob{r}"))
    blobs.append(open(os.path.join(d, f"blob{r}"), "rb").read())
assert rhs.p2p_connect_blobs(rank, blobs)
nr, rk, boxes = rhs.p2p_mailboxes()
assert nr == world and rk == rank and all(boxes), (nr, rk, boxes)
# every rank has mapped the others before anybody's kernel stores into a mailbox
open(os.path.join(d, f"mapped{rank}"), "w").close()
for r in range(world):
    wait_for(os.path.join(d, f"mapped{r}"))

st = rhs.torch_stream()
ops = NVectorOps(0, rhs.stream_ptr, owner=rhs)
rng = np.random.default_rng(100 + rank)
n = 20000 + 7777 * rank                      # ragged: the ranks hold different lengths
n_global = sum(20000 + 7777 * r for r in range(world))
ops.n_global = n_global
out = {"n": np.array([n])}
with torch.cuda.stream(st):
    x = torch.from_numpy(rng.normal(0, 1, n)).cuda()
    y = torch.from_numpy(rng.normal(0, 1, n)).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 2.0, n)).cuda()
    Y = [torch.from_numpy(rng.normal(0, 1, n)).cuda() for _ in range(5)]
    st.synchronize()

    xs = x.clone()
    glo = {"dot": [], "wrms": [], "max": [], "min": [], "multi": []}
    # local values first (no exchange), then the same reductions on the same vectors as global ones; several rounds:
    # both parities of the mailbox slots, sequence numbers running on
    ops.set_peer_allreduce(nr, rk, boxes)
    ops.local(True)
    loc = {"dot": [], "wsq": [], "max": [], "min": [], "multi": []}
    for k in range(3):
        loc["dot"].append(ops.N_VDotProd(x, y)); loc["wsq"].append(ops.N_VWSqrSumLocal(x, w))
        loc["max"].append(ops.N_VMaxNorm(x)); loc["min"].append(ops.N_VMin(x)); loc["multi"].append(ops.N_VDotProdMulti(x, Y))
        x.mul_(1.0 + 0.25 * (k + 1)); st.synchronize()
    ops.local(False)
    x.copy_(xs); st.synchronize()
    for k in range(3):
        glo["dot"].append(ops.N_VDotProd(x, y)); glo["wrms"].append(ops.N_VWrmsNorm(x, w))
        glo["max"].append(ops.N_VMaxNorm(x)); glo["min"].append(ops.N_VMin(x)); glo["multi"].append(ops.N_VDotProdMulti(x, Y))
        x.mul_(1.0 + 0.25 * (k + 1)); st.synchronize()
    # the same mailboxes installed again (a second vector made distributed on this workspace): the sequence goes on
    ops.local(True)
    out["loc_again"] = np.array([ops.N_VDotProd(y, y)])
    ops.local(False)
    ops.set_peer_allreduce(nr, rk, boxes)
    out["glo_again"] = np.array([ops.N_VDotProd(y, y), ops.N_VDotProd(y, y)])
for k_, v_ in loc.items():
    out["loc_" + k_] = np.array(v_, dtype=np.float64)
for k_, v_ in glo.items():
    out["glo_" + k_] = np.array(v_, dtype=np.float64)
np.savez(os.path.join(d, f"result{rank}.tmp.npz"), **out)
os.rename(os.path.join(d, f"result{rank}.tmp.npz"), os.path.join(d, f"result{rank}.npz"))
# nobody unmaps while a peer may still be inside a reduction
for r in range(world):
    wait_for(os.path.join(d, f"result{r}.npz"))
ops.close()
print("rank", rank, "ok")
