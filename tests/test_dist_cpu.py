"""world_size-2 gloo tests (CPU) of the block-cyclic layer: ownership arithmetic and the SUMMA
driver's panel schedule, with a torch stand-in for the local GEMM (the product path uses the DMMA
kernel; see tests/test_dist_gpu.py for the GPU leg)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = """
import sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from lla_b200 import dist as LD
dist.init_process_group("gloo")
world, rank = dist.get_world_size(), dist.get_rank()
for (P, Q) in ((1, 2), (2, 1)):
    grid = LD.make_grid(P, Q)
    m, n, k, nb = 8 * P, 12 * Q, 4 * P * Q, 2
    torch.manual_seed(0)
    A = torch.rand((k, m), dtype=torch.float64)      # column-major images
    B = torch.rand((n, k), dtype=torch.float64)
    la, lb, lc = LD.BlockCyclic(m, k, nb, grid), LD.BlockCyclic(k, n, nb, grid), LD.BlockCyclic(m, n, nb, grid)
    A_loc, B_loc = la.scatter_from_global(A), lb.scatter_from_global(B)
    assert torch.equal(la.gather_to_global(A_loc), A)          # scatter/gather round trip
    C_loc = torch.zeros((n // Q, m // P), dtype=torch.float64)
    for look in (True, False):
        LD.summa(grid, m, n, k, nb, A_loc, B_loc, C_loc, local_gemm=LD.torch_gemm, lookahead=look)
        C = lc.gather_to_global(C_loc)
        want = (A.t() @ B.t()).t()
        assert torch.allclose(C, want, atol=1e-12), (P, Q, look)
    # complex double (configs[4] names ZGEMM): same schedule, panels broadcast as complex tensors
    Az = torch.complex(torch.rand((k, m), dtype=torch.float64), torch.rand((k, m), dtype=torch.float64))
    Bz = torch.complex(torch.rand((n, k), dtype=torch.float64), torch.rand((n, k), dtype=torch.float64))
    Cz_loc = torch.zeros((n // Q, m // P), dtype=torch.complex128)
    LD.summa(grid, m, n, k, nb, la.scatter_from_global(Az), lb.scatter_from_global(Bz), Cz_loc, local_gemm=LD.torch_gemm)
    assert torch.allclose(lc.gather_to_global(Cz_loc), (Az.t() @ Bz.t()).t(), atol=1e-12), (P, Q, "complex")
assert LD.choose_grid(8) == (2, 4) and LD.choose_grid(4) == (2, 2) and LD.choose_grid(2) == (1, 2)
dist.destroy_process_group()
"""


def test_summa_block_cyclic_two_ranks(tmp_path):
    script = tmp_path / "summa.py"
    script.write_text(SCRIPT.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]


CHOL_SCRIPT = """
import sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from lla_b200 import dist as LD
dist.init_process_group("gloo")
world, rank = dist.get_world_size(), dist.get_rank()
ops = LD.TorchOps()
for (P, Q) in ((1, 2), (2, 1)):
    grid = LD.make_grid(P, Q)
    nb = 3
    n = nb * P * Q * 4
    torch.manual_seed(1)
    G = torch.rand((n, n), dtype=torch.float64)
    S = G @ G.t() + n * torch.eye(n, dtype=torch.float64)
    # what LLA hands to POTRF('U'): only the upper triangle of the column-major view is meaningful
    full = torch.triu(S) + torch.tril(torch.full((n, n), 777.0, dtype=torch.float64), -1)
    lay = LD.BlockCyclic(n, n, nb, grid)
    A_loc = lay.scatter_from_global(full.t().contiguous())
    info = LD.pdpotrf(grid, n, nb, A_loc, ops)
    assert info == 0
    out = lay.gather_to_global(A_loc).t()
    U = torch.triu(out)
    want = torch.linalg.cholesky(S).t()
    assert torch.allclose(U, want, atol=1e-11), (P, Q, (U - want).abs().max())
    assert torch.equal(torch.tril(out, -1), torch.tril(full, -1)), "lower triangle must not be touched"
    # solve with the distributed factor, right-hand sides replicated
    b = torch.rand((n, 5), dtype=torch.float64)
    B = b.t().contiguous().clone()
    LD.pdpotrs(grid, n, nb, A_loc, B, ops)
    assert torch.allclose(S @ B.t(), b, atol=1e-10), (P, Q)
    # not positive definite: INFO = index of the first failing leading minor, the same on every rank
    bad = S.clone(); kbad = nb * 2 + 1; bad[kbad, kbad] = -1.0
    A_loc = lay.scatter_from_global(bad.t().contiguous())
    info = LD.pdpotrf(grid, n, nb, A_loc, ops)
    assert info == kbad + 1, (info, kbad + 1)
dist.destroy_process_group()
"""


def test_block_cyclic_cholesky_two_ranks(tmp_path):
    script = tmp_path / "pchol.py"
    script.write_text(CHOL_SCRIPT.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29543", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]


LU_SCRIPT = """
import sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from lla_b200 import dist as LD
dist.init_process_group("gloo")
world, rank = dist.get_world_size(), dist.get_rank()
ops = LD.TorchOps()
grid = LD.make_grid(1, 2)
nb = 3
n = nb * 2 * 4
torch.manual_seed(2)
M = torch.rand((n, n), dtype=torch.float64) * 10            # as lu-random-test, tests/linear-algebra.lisp:112
lay = LD.BlockCyclic(n, n, nb, grid)
want_lu, want_piv = torch.linalg.lu_factor(M)
for look in (True, False):
    A_loc = lay.scatter_from_global(M.t().contiguous())
    ipiv = torch.zeros(n, dtype=torch.int32)
    info = LD.pdgetrf(grid, n, nb, A_loc, ipiv, ops, lookahead=look)
    assert info == 0
    assert torch.equal(ipiv, want_piv.to(torch.int32)), "ipiv must not depend on the process count"
    got = lay.gather_to_global(A_loc).t()
    assert torch.allclose(got, want_lu, atol=1e-11)
    b = torch.rand((n, 5), dtype=torch.float64)
    B = b.t().contiguous().clone()
    LD.pdgetrs(grid, n, nb, A_loc, ipiv, B, ops)
    assert torch.allclose(M @ B.t(), b, atol=1e-9)
# exactly singular: INFO = first zero pivot, factorisation completes (lu-singular, tests/linear-algebra.lisp:106-107)
S = M.clone(); S[:, 4] = 0.0
A_loc = lay.scatter_from_global(S.t().contiguous())
ipiv = torch.zeros(n, dtype=torch.int32)
assert LD.pdgetrf(grid, n, nb, A_loc, ipiv, ops) == 5
try:
    LD.pdgetrf(LD.make_grid(2, 1), n, nb, A_loc, ipiv, ops)
    raise SystemExit("a 2 x 1 grid must be refused")
except ValueError:
    pass
dist.destroy_process_group()
"""


def test_block_column_lu_two_ranks(tmp_path):
    script = tmp_path / "plu.py"
    script.write_text(LU_SCRIPT.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29545", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]


GRID22_SCRIPT = """
import sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from lla_b200 import dist as LD
dist.init_process_group("gloo")
ops = LD.TorchOps()
grid = LD.make_grid(2, 2)
nb = 2
n = nb * 2 * 5                     # 10 blocks per dimension: every process owns diagonal and off-diagonal blocks
torch.manual_seed(3)
G = torch.rand((n, n), dtype=torch.float64)
S = G @ G.t() + n * torch.eye(n, dtype=torch.float64)
lay = LD.BlockCyclic(n, n, nb, grid)
for look in (True, False):
    A_loc = lay.scatter_from_global(torch.triu(S).t().contiguous())
    assert LD.pdpotrf(grid, n, nb, A_loc, ops, lookahead=look) == 0
    U = torch.triu(lay.gather_to_global(A_loc).t())
    assert torch.allclose(U, torch.linalg.cholesky(S).t(), atol=1e-11), look
b = torch.rand((n, 3), dtype=torch.float64)
B = b.t().contiguous().clone()
LD.pdpotrs(grid, n, nb, A_loc, B, ops)
assert torch.allclose(S @ B.t(), b, atol=1e-10)
# SUMMA on the same 2 x 2 grid
A = torch.rand((n, n), dtype=torch.float64); Bm = torch.rand((n, n), dtype=torch.float64)
C_loc = torch.zeros((n // 2, n // 2), dtype=torch.float64)
LD.summa(grid, n, n, n, nb, lay.scatter_from_global(A), lay.scatter_from_global(Bm), C_loc, local_gemm=LD.torch_gemm)
assert torch.allclose(lay.gather_to_global(C_loc), (A.t() @ Bm.t()).t(), atol=1e-12)
dist.destroy_process_group()
"""


def test_two_by_two_grid_four_ranks(tmp_path):
    script = tmp_path / "grid22.py"
    script.write_text(GRID22_SCRIPT.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4",
                        "--master-addr", "127.0.0.1", "--master-port", "29547", str(script)],
                       capture_output=True, text=True, env=env, timeout=400)
    assert r.returncode == 0, r.stderr[-3000:]
