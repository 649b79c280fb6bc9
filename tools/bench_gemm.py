"""Times the latent-side GEMM shapes of cfg 1 through hn_op_gemm (run on the GPU box)."""
import ctypes, sys, os, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import healnet_b200
lib = healnet_b200.load_library()
st = torch.cuda.current_stream().cuda_stream
RM = int(os.environ.get("HN_BENCH_ROWS", "2048"))   # latent rows = batch * l_c
shapes = [("Q' small", 2048, 256, 512, 0, 0), ("Q tab", 2048, 512, 512, 0, 1), ("out-proj", 2048, 512, 512, 3, 0),
          ("FF1 gate", 2048, 4096, 512, 1, 1), ("FF2 res", 2048, 512, 2048, 2, 0), ("QKV self", 2048, 1536, 512, 0, 1),
          ("KV tab", 4, 1024, 2005, 0, 1)]
tot = 0.0
counts = {"Q' small": 6, "Q tab": 3, "out-proj": 18, "FF1 gate": 18, "FF2 res": 18, "QKV self": 9, "KV tab": 3}
shapes = [(n, RM if M == 2048 else M, N, K, e, so) for (n, M, N, K, e, so) in shapes]
for name, M, N, K, epi, split_out in shapes:
    seg = (K + 63) // 64 * 64
    A = (torch.randn(M, 2 * seg, device="cuda") * 0.1).half()
    B = (torch.randn(N, 2 * seg, device="cuda") * 0.1).half()
    bias = torch.randn(N, device="cuda")
    half_out = epi in (0, 1)
    ncols = N // 2 if epi == 1 else N
    out = torch.zeros(M, 2 * ncols if half_out else N, device="cuda", dtype=torch.float16 if half_out else torch.float32)
    ldo = 2 * ncols if half_out else N
    args = (A.data_ptr(), B.data_ptr(), M, N, K, 2 * seg, 2 * seg, epi, 0, bias.data_ptr(), out.data_ptr(), ldo, 3, seg,
            seg, ncols if (half_out and split_out) else 0, st)
    for _ in range(5):
        assert lib.hn_op_gemm(*args) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 50
    for _ in range(n):
        lib.hn_op_gemm(*args)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    fl = 2.0 * M * N * K * 3
    tot += us * counts[name]
    print(f"{name:10s} M={M:5d} N={N:5d} K={K:5d}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s   x{counts[name]} = {us * counts[name] / 1e3:.3f} ms")
print(f"per forward: {tot / 1e3:.3f} ms")
