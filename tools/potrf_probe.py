"""Times abo_potrf_dev (blocked FP64 Cholesky) at a few sizes; used under ncu for launch lists."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import abo_b200 as abo

ctx = abo.default_context(0)
st = torch.cuda.ExternalStream(ctx.stream())
sizes = [int(a) for a in sys.argv[1:]] or [8192]
for n in sizes:
    g = torch.Generator(device="cuda").manual_seed(n)
    X = torch.rand((n, 20), dtype=torch.float64, device="cuda", generator=g)
    K0 = torch.exp(-0.5 * torch.cdist(X, X) ** 2) + 1e-2 * torch.eye(n, dtype=torch.float64, device="cuda")
    A = torch.empty_like(K0)
    best = 1e9
    for it in range(3):
        A.copy_(K0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); ctx.potrf_dev(A.data_ptr(), n, n); e1.record(st); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fl = n ** 3 / 3 + n ** 2 / 2
    L = torch.tril(A)
    err = (torch.linalg.norm(L @ L.T - K0) / torch.linalg.norm(K0)).item()
    ck = ctx.potf2_clocks()        # stamps of CTA 0 of the last potf2 launch: 0 start, 1 staged, 2 factor group done, 6 end, 11-14 last inverse round
    print("   potf2 (cycles): staged", ck[1] - ck[0], "| factor group done", ck[2] - ck[0], "| inverse round 3: X_33", ck[13] - ck[12],
          "multiply+store", ck[14] - ck[13], "| kernel end", ck[6] - ck[0])
    print("   diag 0 done", ck[4] - ck[0], "| warp 1 released", ck[15] - ck[0], "| substituted", ck[7] - ck[0], "| panel barrier", ck[8] - ck[0],
          "| critical tile", ck[10] - ck[0], "| diag 1 starts", ck[5] - ck[0], "| diag 1 done", ck[9] - ck[0])
    print(f"n={n} potrf {best:.3f} ms  {fl / best / 1e9:.2f} TFLOP/s  relres={err:.2e}", flush=True)
