"""Timeline of the small-context attention kernel (debug): CTA 0 stamps clock64() at a few points of 48 tiles for
the first softmax warp and the UMMA issuer of every row block (HN_TR in xattn_small.cu). Prints, per variant / mode,
the mean clocks between consecutive points and the period per tile. Run on the GPU box:
    HN_SMALL_VARIANT=0 HN_POLY_MODE=5 python tools/trace_attn.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import healnet_b200  # noqa: E402
from healnet_b200 import HealNet  # noqa: E402

MAXG, NT, NP = 4, 48, 8


def main():
    lib = healnet_b200.load_library()
    lib.hn_debug_set_trace.restype = None
    lib.hn_debug_set_trace.argtypes = [ctypes.c_void_p]
    torch.manual_seed(0)
    model = HealNet(n_modalities=1, channel_dims=[3], num_spatial_axes=[3], out_dims=4, l_c=512, l_d=512, depth=1).eval().cuda()
    x = [torch.rand(4, 12, 224, 224, 3, device="cuda")]
    model(x)
    torch.cuda.synchronize()
    buf = torch.zeros(2 * MAXG * NT * NP, dtype=torch.int64, device="cuda")
    lib.hn_debug_set_trace(buf.data_ptr())
    model(x)
    torch.cuda.synchronize()
    lib.hn_debug_set_trace(None)
    t = buf.cpu().view(2 * MAXG, NT, NP).double()
    print("variant", os.environ.get("HN_SMALL_VARIANT", "default"), "mode", os.environ.get("HN_POLY_MODE", "default"))
    for slot in range(2 * MAXG):
        rec = t[slot]
        if rec.abs().sum() == 0:
            continue
        used = [k for k in range(NP) if rec[:, k].abs().sum() > 0]
        used.sort(key=lambda k: (rec[:, k] - rec[:, 0]).mean().item())
        period = (rec[-1, used[0]] - rec[0, used[0]]) / (NT - 1)
        segs = []
        for a, b in zip(used[:-1], used[1:]):
            segs.append("p%d->p%d %6.0f" % (a, b, (rec[:, b] - rec[:, a]).mean().item()))
        # tail: last point of tile i -> first point of tile i+1
        segs.append("p%d->next p%d %6.0f" % (used[-1], used[0], (rec[1:, used[0]] - rec[:-1, used[-1]]).mean().item()))
        kind = "softmax g%d" % slot if slot < MAXG else "issuer  g%d" % (slot - MAXG)
        print("%s: period/tile %7.0f clk | %s" % (kind, period.item(), " | ".join(segs)))
    # relative phase of the groups (first point)
    base = t[0, :, 0]
    for slot in range(1, MAXG):
        if t[slot].abs().sum() > 0:
            print("  softmax g%d starts its tiles %+.0f clk after g0 (mean)" % (slot, (t[slot, :, 0] - base).mean().item()))


if __name__ == "__main__":
    main()
