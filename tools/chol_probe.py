import sys, os, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0)
side = len(sys.argv) > 1 and sys.argv[1] == "side"
st = torch.cuda.Stream() if side else torch.cuda.current_stream()
with torch.cuda.stream(st):
    D.use_torch_stream()
    for n in (8192, 16384):
        G = torch.rand((n, n), dtype=torch.float64, device="cuda")
        S0 = G @ G.t() + n * torch.eye(n, dtype=torch.float64, device="cuda"); del G
        A = S0.clone(); info = torch.zeros(1, dtype=torch.int32, device="cuda")
        best = 1e9
        for _ in range(3):
            A.copy_(S0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); D.dpotrf("U", n, A.data_ptr(), n, info.data_ptr()); e1.record(st); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{'side' if side else 'default'} stream, NO_LOOKAHEAD={os.environ.get('LLACUDA_NO_LOOKAHEAD')}: dpotrf n={n}: {best:.2f} ms {n**3/3/best/1e9:.2f} TF", flush=True)
        del A, S0
