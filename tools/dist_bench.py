"""Multi-GPU legs (run under torchrun, one rank per GPU):
  * SUMMA DGEMM on a 2-D block-cyclic grid with NCCL panel broadcasts (configs[4] shape),
  * batched 32x32 solves sharded one slice per GPU (configs[3]).
Prints one JSON line per leg on rank 0.  Timing: CUDA events, max over ranks.
  torchrun --nproc-per-node N tools/dist_bench.py [--size 32768] [--block 1024] [--batch 200000]
"""
import argparse, json, os, sys, math
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lla_b200 import _lib, device as D, dist as LD, batched as BT

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=16384)
ap.add_argument("--block", type=int, default=1024)
ap.add_argument("--batch", type=int, default=200000)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
_lib.init(lr); D.use_torch_stream()
dev = torch.device("cuda", lr)
grid = LD.make_grid()
n, nb = args.size, args.block
P, Q = grid.P, grid.Q

def elem(i, j, salt):
    # rank-independent synthetic entries so that any rank can regenerate any row / column
    return torch.sin(i * 0.37 + j * 0.11 + salt) + 0.25 * torch.cos(i * 0.05 - j * 0.21)

def local_image(layout, salt):
    nbk = layout.nb
    rows = torch.cat([torch.arange(I * nbk, (I + 1) * nbk) for I in layout.local_block_rows()]).to(dev).double()
    cols = torch.cat([torch.arange(J * nbk, (J + 1) * nbk) for J in layout.local_block_cols()]).to(dev).double()
    return elem(rows[None, :], cols[:, None], salt).contiguous(), rows, cols   # (nloc, mloc) image

la, lb, lc = LD.BlockCyclic(n, n, nb, grid), LD.BlockCyclic(n, n, nb, grid), LD.BlockCyclic(n, n, nb, grid)
A_loc, _, _ = local_image(la, 0.0)
B_loc, _, _ = local_image(lb, 1.0)
C_loc = torch.zeros((n // Q, n // P), dtype=torch.float64, device=dev)
_, crow, ccol = local_image(lc, 0.0)

def timed(fn):
    best = 1e30
    for _ in range(args.reps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, t.item())
    return best

LD.summa(grid, n, n, n, nb, A_loc, B_loc, C_loc)          # warm-up
ms = timed(lambda: LD.summa(grid, n, n, n, nb, A_loc, B_loc, C_loc))
# spot check 32 local entries against regenerated rows / columns
g = torch.Generator(device="cpu").manual_seed(rank)
ii = torch.randint(0, crow.numel(), (32,), generator=g); jj = torch.randint(0, ccol.numel(), (32,), generator=g)
kk = torch.arange(n, device=dev).double()
err = 0.0
for a, b in zip(ii.tolist(), jj.tolist()):
    gi, gj = crow[a], ccol[b]
    want = (elem(gi, kk, 0.0) * elem(kk, gj, 1.0)).sum()
    err = max(err, abs((C_loc[b, a] - want).item()))
t = torch.tensor([err], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"leg": "summa_dgemm", "n": n, "nb": nb, "grid": [P, Q], "n_gpus": world, "ms": ms,
                      "tflops": 2.0 * n ** 3 / ms / 1e9, "max_abs_err_sampled": t.item(),
                      "err_bound_k_eps": 4 * math.sqrt(n) * 2.2e-16 * n * 1.6}), flush=True)
del A_loc, B_loc, C_loc

# ---- batched small solves, one contiguous slice per GPU, no collective on the data path
lo, hi = BT.shard_range(args.batch, rank, world)
mine = hi - lo
gen = torch.Generator(device=dev).manual_seed(4 + rank)
A0 = torch.rand((mine, 32, 32), dtype=torch.float64, device=dev, generator=gen) * 10
B0 = torch.rand((mine, 32), dtype=torch.float64, device=dev, generator=gen) * 10
A, B = A0.clone(), B0.clone()
info = torch.zeros(mine, dtype=torch.int32, device=dev)
def run():
    D.dgesv_batched(32, 1, mine, A.data_ptr(), 1024, None, B.data_ptr(), 32, info.data_ptr(), 0)
best = 1e30
for _ in range(args.reps + 1):
    A.copy_(A0); B.copy_(B0); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = min(best, t.item())
x = B.unsqueeze(2); Am = A0.transpose(1, 2)
res = (torch.linalg.norm(Am @ x - B0.unsqueeze(2), dim=(1, 2)) / (32 * 2.22e-16 * torch.linalg.norm(Am, dim=(1, 2)) * torch.linalg.norm(x, dim=(1, 2)))).max()
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    byts = args.batch * (1024 * 8 + 32 * 16 + 4)
    print(json.dumps({"leg": "batched_dgesv_32", "batch": args.batch, "n_gpus": world, "ms": best,
                      "systems_per_s": args.batch / best * 1e3, "algorithmic_GBps": byts / best / 1e6,
                      "max_scaled_residual": res.item()}), flush=True)
dist.destroy_process_group()
