"""Block-cyclic Cholesky + solve on N GPUs (configs[2]): torchrun --nproc-per-node N tools/dist_chol_bench.py --size 32768 --block 1024
One JSON line on rank 0.  Timing: CUDA events, max over ranks, best of --reps."""
import argparse, json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lla_b200 import _lib, device as D, dist as LD

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=16384)
ap.add_argument("--block", type=int, default=1024)
ap.add_argument("--nrhs", type=int, default=64)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--lookahead", type=int, default=1)
args = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
_lib.init(lr); D.use_torch_stream()
dev = torch.device("cuda", lr)
grid = LD.make_grid()
n, nb, P, Q = args.size, args.block, grid.P, grid.Q
lay = LD.BlockCyclic(n, n, nb, grid)

def elem(i, j):
    # symmetric, rank-independent synthetic entries; + 2.6 n on the diagonal makes the matrix SPD (Gershgorin)
    return torch.sin((i + j) * 0.37) + 0.25 * torch.cos((i - j).abs() * 0.05) + torch.where(i == j, 2.6 * n, 0.0)

rows = torch.cat([torch.arange(I * nb, (I + 1) * nb) for I in lay.local_block_rows()]).to(dev).double()
cols = torch.cat([torch.arange(J * nb, (J + 1) * nb) for J in lay.local_block_cols()]).to(dev).double()
A0 = elem(rows[None, :], cols[:, None]).contiguous()      # (nloc, mloc) image
A_loc = A0.clone()
# right-hand side b = A x0 with x0 = 1: row sums, each rank a slice of rows, then all_gather
lo, hi = rank * n // world, (rank + 1) * n // world
ii = torch.arange(lo, hi, device=dev).double()
jj = torch.arange(n, device=dev).double()
bs = torch.cat([elem(ii[c:c + 512, None], jj[None, :]).sum(dim=1) for c in range(0, hi - lo, 512)])
parts = [torch.empty_like(bs) for _ in range(world)]
dist.all_gather(parts, bs)
b = torch.cat(parts)
B0 = b[None, :].repeat(args.nrhs, 1).contiguous()           # (nrhs, n) image

def timed(fn, prep):
    best = 1e30
    for _ in range(args.reps):
        prep(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, t.item())
    return best

infos = []
kw = {"lookahead": bool(args.lookahead)} if "lookahead" in LD.pdpotrf.__code__.co_varnames else {}
LD.pdpotrf(grid, n, nb, A_loc, **kw)                          # warm-up (NCCL channels, workspaces)
ms_f = timed(lambda: infos.append(LD.pdpotrf(grid, n, nb, A_loc, **kw)), lambda: A_loc.copy_(A0))
B = B0.clone()
ms_s = timed(lambda: LD.pdpotrs(grid, n, nb, A_loc, B), lambda: B.copy_(B0))
err = (B - 1.0).abs().max()
dist.all_reduce(err, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"leg": "pdpotrf+pdpotrs", "n": n, "nb": nb, "grid": [P, Q], "n_gpus": world, "nrhs": args.nrhs,
                      "potrf_ms": ms_f, "potrf_tflops": n ** 3 / 3 / ms_f / 1e9, "potrs_ms": ms_s, "info": infos[-1],
                      "max_abs_err_x_vs_ones": err.item(), "lookahead": bool(args.lookahead)}), flush=True)
dist.destroy_process_group()
