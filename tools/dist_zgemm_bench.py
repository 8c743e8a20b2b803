"""SUMMA ZGEMM on N GPUs (configs[4], complex double): torchrun --nproc-per-node N tools/dist_zgemm_bench.py --size 32768 --block 1024
One JSON line on rank 0.  Timing: CUDA events, max over ranks.  32 sampled entries per rank checked against regenerated rows / columns."""
import argparse, json, os, sys, math
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lla_b200 import _lib, device as D, dist as LD

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--block", type=int, default=1024)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
_lib.init(lr); D.use_torch_stream()
dev = torch.device("cuda", lr)
grid = LD.make_grid()
n, nb, P, Q = args.size, args.block, grid.P, grid.Q

def elem(i, j, salt):
    re = torch.sin(i * 0.37 + j * 0.11 + salt) + 0.25 * torch.cos(i * 0.05 - j * 0.21)
    im = torch.cos(i * 0.23 - j * 0.13 + salt) * 0.5
    return torch.complex(re, im)

lay = LD.BlockCyclic(n, n, nb, grid)
rows = torch.cat([torch.arange(I * nb, (I + 1) * nb) for I in lay.local_block_rows()]).to(dev).double()
cols = torch.cat([torch.arange(J * nb, (J + 1) * nb) for J in lay.local_block_cols()]).to(dev).double()
A_loc = elem(rows[None, :], cols[:, None], 0.0).contiguous()
B_loc = elem(rows[None, :], cols[:, None], 1.0).contiguous()
C_loc = torch.zeros((n // Q, n // P), dtype=torch.complex128, device=dev)
best = 1e30
for it in range(args.reps + 1):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); LD.summa(grid, n, n, n, nb, A_loc, B_loc, C_loc); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if it > 0: best = min(best, t.item())
g = torch.Generator(device="cpu").manual_seed(rank)
ii = torch.randint(0, rows.numel(), (32,), generator=g); jj = torch.randint(0, cols.numel(), (32,), generator=g)
kk = torch.arange(n, device=dev).double()
err = torch.zeros(1, device=dev, dtype=torch.float64)
for a, b in zip(ii.tolist(), jj.tolist()):
    want = (elem(rows[a], kk, 0.0) * elem(kk, cols[b], 1.0)).sum()
    err = torch.maximum(err, (C_loc[b, a] - want).abs().reshape(1))
dist.all_reduce(err, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"leg": "summa_zgemm", "n": n, "nb": nb, "grid": [P, Q], "n_gpus": world, "ms": best,
                      "real_tflops": 8.0 * n ** 3 / best / 1e9, "max_abs_err_sampled": err.item(),
                      "err_bound": 8 * math.sqrt(n) * 2.2e-16 * n * 2.0}), flush=True)
dist.destroy_process_group()
