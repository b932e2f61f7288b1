import torch, json
torch.backends.cuda.matmul.allow_tf32 = False
for (M, N) in [(128, 4_000_000), (64, 8_000_000), (32, 16_000_000)]:
    k = torch.randn(N, M, dtype=torch.float64, device='cuda')
    B = torch.randn(M, M, dtype=torch.float64, device='cuda')
    g = torch.rand(N, dtype=torch.float64, device='cuda')
    def f1(): return k @ B              # N x M x M
    def f2(): return (k * g[:, None]).t() @ k
    for name, f, fl in [("kappa@B", f1, 2.0*N*M*M), ("(kappa*g)^T@kappa", f2, 2.0*N*M*M)]:
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"op": name, "M": M, "N": N, "ms": best, "tflops": fl / best * 1e-9}))
