"""Launcher-mode check of the multi-GPU engine (run under torchrun, one rank per GPU; tests/test_pdist_gpu.py spawns it).
Every rank builds its block-cyclic part of seeded global matrices, calls the llacuda_*_local rank functions, the parts
are gathered on rank 0 and compared with NumPy / the single-GPU library.  torch.distributed is plumbing only."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lla_b200 import _lib, lla as L, pdist as PD  # noqa: E402

EPS = np.finfo(np.float64).eps


def gather_global(bc, loc_img, rank, world):
    """inverse of BlockCyclic.scatter on rank 0 (others return None); loc_img: (nloc, mloc) torch CUDA tensor"""
    parts = [torch.empty_like(loc_img) for _ in range(world)] if rank == 0 else None
    dist.gather(loc_img, parts, dst=0)
    if rank != 0:
        return None
    full = np.empty((bc.Np, bc.Mp))
    for r, part in enumerate(parts):
        o = PD.BlockCyclic(bc.Mp, bc.Np, bc.nb, bc.P, bc.Q, r)
        full[np.ix_(o.local_cols(), o.local_rows())] = part.cpu().numpy()
    return full   # column-major image: full[j, i] = A(i, j)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.init(local)
    dev = torch.device("cuda", local)
    PD.init_from_torch()
    grids = [(p, world // p) for p in range(1, world + 1) if world % p == 0]
    rng = np.random.default_rng(7)   # same stream on every rank: every rank knows the global matrices
    nb = 128
    for (P, Q) in grids:
        PD.set_grid(P, Q)
        n = nb * int(PD.lcm(P, Q)) * 3
        bc = PD.BlockCyclic(n, n, nb, P, Q, rank)
        # ---- SUMMA
        a, b = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
        A = torch.as_tensor(bc.scatter(np.ascontiguousarray(a.T)), device=dev)
        B = torch.as_tensor(bc.scatter(np.ascontiguousarray(b.T)), device=dev)
        C = torch.zeros_like(A)
        torch.cuda.synchronize()
        PD.pdgemm_local(n, n, n, nb, 1.0, A.data_ptr(), bc.mloc, B.data_ptr(), bc.mloc, 0.0, C.data_ptr(), bc.mloc)
        full = gather_global(bc, C, rank, world)
        if rank == 0:
            assert np.max(np.abs(full.T - a @ b)) <= 8 * np.sqrt(n) * EPS * np.abs(a @ b).max() * 4, ("pdgemm", P, Q)
        # ---- LU + solve
        a = rng.uniform(0, 10, (n, n))
        rhs = rng.uniform(0, 10, (n, 3))
        A = torch.as_tensor(bc.scatter(np.ascontiguousarray(a.T)), device=dev)
        ipiv = torch.zeros(n, dtype=torch.int32, device=dev)
        Bd = torch.as_tensor(np.ascontiguousarray(rhs.T), device=dev)      # (nrhs, n) = column-major n x nrhs, replicated
        torch.cuda.synchronize()
        info = PD.pdgetrf_local(n, nb, A.data_ptr(), bc.mloc, ipiv.data_ptr())
        assert info == 0
        PD.pdgetrs_local(n, nb, A.data_ptr(), bc.mloc, ipiv.data_ptr(), Bd.data_ptr(), n, 3)
        torch.cuda.synchronize()
        full = gather_global(bc, A, rank, world)
        if rank == 0:
            one = L.lu(a)          # single-GPU factorisation of the same matrix
            assert np.array_equal(ipiv.cpu().numpy(), one.ipiv_internal), ("pdgetrf ipiv", P, Q)
            assert np.max(np.abs(full.T - one.lu)) <= 50 * n * EPS * np.abs(one.lu).max(), ("pdgetrf lu", P, Q)
            x = Bd.cpu().numpy().T
            res = np.linalg.norm(a @ x - rhs) / (n * EPS * np.linalg.norm(a) * np.linalg.norm(x))
            assert res <= 10, ("pdgetrs", P, Q, res)
        # ---- the same system through the one-call entry point (right-hand sides ride along on block columns)
        A2 = torch.as_tensor(bc.scatter(np.ascontiguousarray(a.T)), device=dev)
        ipiv2 = torch.zeros(n, dtype=torch.int32, device=dev)
        B2 = torch.as_tensor(np.ascontiguousarray(rhs.T), device=dev)
        torch.cuda.synchronize()
        info = PD.pdgesv_local(n, nb, A2.data_ptr(), bc.mloc, ipiv2.data_ptr(), B2.data_ptr(), n, 3)
        torch.cuda.synchronize()
        assert info == 0
        assert torch.equal(ipiv2, ipiv) and torch.equal(A2, A), ("pdgesv_local factors", P, Q)
        if rank == 0:
            x2 = B2.cpu().numpy().T
            res = np.linalg.norm(a @ x2 - rhs) / (n * EPS * np.linalg.norm(a) * np.linalg.norm(x2))
            assert res <= 10, ("pdgesv_local", P, Q, res)
            assert np.max(np.abs(x2 - x)) <= 1e4 * n * EPS * np.abs(x).max(), ("pdgesv_local vs pdgetrs", P, Q)
        # ---- Cholesky + solve
        g = rng.uniform(0, 1, (n, n))
        s = g @ g.T + n * np.eye(n)
        A = torch.as_tensor(bc.scatter(np.ascontiguousarray(s.T)), device=dev)
        Bd = torch.as_tensor(np.ascontiguousarray(rhs.T), device=dev)
        torch.cuda.synchronize()
        info = PD.pdpotrf_local(n, nb, A.data_ptr(), bc.mloc)
        assert info == 0
        PD.pdpotrs_local(n, nb, A.data_ptr(), bc.mloc, Bd.data_ptr(), n, 3)
        torch.cuda.synchronize()
        full = gather_global(bc, A, rank, world)
        if rank == 0:
            u = np.triu(full.T)
            assert np.max(np.abs(u.T @ u - s)) <= 20 * n * EPS * np.abs(s).max(), ("pdpotrf", P, Q)
            x = Bd.cpu().numpy().T
            res = np.linalg.norm(s @ x - rhs) / (n * EPS * np.linalg.norm(s) * np.linalg.norm(x))
            assert res <= 10, ("pdpotrs", P, Q, res)
        if rank == 0:
            print(f"grid {P}x{Q}: pdgemm / pdgetrf+pdgetrs / pdpotrf+pdpotrs ok (n={n}, nb={nb})", flush=True)
    dist.barrier()
    PD.finalize()
    if rank == 0:
        print("pdist_local_check ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
