"""GPU probe: validates the UMMA / TMA / TMEM conventions of tc05.cuh on a real B200 (run under gpurun)."""
import ctypes, os, sys, itertools
import torch

lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "..", "healnet_b200", "libhealnet_b200.so"))
lib.hn_debug_probe.restype = ctypes.c_int
lib.hn_debug_probe.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 2 + [ctypes.c_void_p] * 4


def run(kd, vd, ov=None, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    Q = (torch.randn(128, kd, generator=g) * 0.5).half().cuda()
    K = (torch.randn(64, kd, generator=g) * 0.5).half().cuda()
    V = (torch.randn(64, vd, generator=g) * 0.5).half().cuda()
    S = torch.full((128, 64), float("nan"), device="cuda")
    U = torch.full((128, vd), float("nan"), device="cuda")
    ovp = None
    if ov is not None:
        arr = (ctypes.c_int * 10)(*ov)
        ovp = ctypes.cast(arr, ctypes.c_void_p)
    rc = lib.hn_debug_probe(Q.data_ptr(), K.data_ptr(), V.data_ptr(), kd, vd, S.data_ptr(), U.data_ptr(), ovp,
                            torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    Sref = Q.float() @ K.float().t()
    Uref = Sref.half().float() @ V.float()
    es = (S - Sref).abs().max().item()
    eu = (U - Uref).abs().max().item()
    return rc, es, eu


print(torch.cuda.get_device_name(0), torch.version.cuda)
for kd, vd in [(64, 64), (32, 32), (32, 64), (64, 32)]:
    rc, es, eu = run(kd, vd)
    print(f"default kd={kd} vd={vd}: rc={rc} errS={es:.3e} errU={eu:.3e}", flush=True)
    if not (es < 1e-2):
        for lbo, sbo in itertools.product([0, 16, 64, 128, 512, 1024], [256, 512, 1024]):
            ov = [lbo, sbo, lbo, sbo, -1, -1, -1, -1, -1, -1]
            rc, es2, _ = run(kd, vd, ov)
            if es2 < 1e-2:
                print(f"   S ok with q/k lbo={lbo} sbo={sbo}")
    if not (eu < 1e-2):
        for lbo, sbo, kadv in itertools.product([0, 16, 64, 128, 512, 1024, 2048], [64, 128, 256, 512, 1024, 2048],
                                                [512, 1024, 2048]):
            ov = [-1, -1, -1, -1, lbo, sbo, -1, kadv, -1, -1]
            rc, _, eu2 = run(kd, vd, ov)
            if eu2 < 1e-2:
                print(f"   U ok with v lbo={lbo} sbo={sbo} kadv={kadv}")
