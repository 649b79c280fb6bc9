"""GPU (B200): size-independent properties at BASELINE.json's full token counts, where the oracle would take
minutes and tens of GB. Softmax attention over a set of tokens offers:
  * replication invariance: with Fourier features off, N identical tokens attend exactly like 1 token;
  * mask == removal: masking tokens out equals deleting them (set semantics without positional features);
  * permutation invariance over the token axis (no positional features);
and the path as a whole is per-sample (batch independence at full size)."""
import pytest
import torch

from healnet_b200 import HealNet

pytestmark = pytest.mark.gpu


def _m(**kw):
    torch.manual_seed(0)
    base = dict(n_modalities=1, channel_dims=[3], num_spatial_axes=[3], out_dims=4, l_c=512, l_d=512,
                fourier_encode_data=False)
    base.update(kw)
    return HealNet(**base).eval().cuda()


def test_replication_invariance_full_volume():
    """602 112 identical voxels (cfg 1 volume extent 12x224x224) == a single voxel."""
    m = _m()
    tok = torch.rand(2, 1, 1, 1, 3, device="cuda")
    big = tok.expand(2, 12, 224, 224, 3).contiguous()
    torch.testing.assert_close(m([big]), m([tok]), rtol=1e-3, atol=1e-4)


def test_mask_equals_removal_and_permutation_full_wsi():
    """cfg 4 WSI extent: 8192 tokens x 768 features (generic K/V projection path)."""
    m = _m(channel_dims=[768], num_spatial_axes=[1], depth=1)
    x = torch.rand(2, 8192, 768, device="cuda")
    keep = torch.rand(2, 8192, device="cuda") > 0.5
    keep[1] = keep[0]  # same kept count per sample so the removed version is rectangular
    masked = m([x], mask=keep)
    removed = m([x[:, keep[0]]])
    torch.testing.assert_close(masked, removed, rtol=1e-3, atol=1e-4)
    perm = torch.randperm(8192, device="cuda")
    torch.testing.assert_close(m([x[:, perm]]), m([x]), rtol=1e-3, atol=1e-4)


def test_batch_independence_cfg1_full_size():
    """BASELINE config 1 shapes (tab 1x2000, img 224x224x3, vol 12x224x224x3, latent 512x512), batch 2 vs 1+1."""
    torch.manual_seed(0)
    m = HealNet(n_modalities=3, channel_dims=[2000, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, l_c=512,
                l_d=512).eval().cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [torch.rand(2, 1, 2000, device="cuda", generator=g), torch.rand(2, 224, 224, 3, device="cuda", generator=g),
          torch.rand(2, 12, 224, 224, 3, device="cuda", generator=g)]
    full = m(xs)
    assert bool(torch.isfinite(full).all())
    for i in range(2):
        torch.testing.assert_close(m([t[i:i + 1] for t in xs]), full[i:i + 1], rtol=1e-3, atol=1e-4)


def test_pinned_host_inputs_in_flight_match_device_inputs():
    """Serving loop on pinned HOST inputs: several forwards in flight (the module stages each call's inputs in one of
    `host_staging_depth` persistent device buffer sets, copied on a side stream that waits only for the forward that
    last read the set). Every call must see ITS inputs: results equal the forwards of the same tensors passed as device
    tensors, bit for bit, also when the queue is deeper than the ring and when shapes change between calls."""
    torch.manual_seed(0)
    kw = dict(n_modalities=3, channel_dims=[200, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, l_c=128, l_d=128, depth=2)
    m = HealNet(**kw).eval().cuda()
    g = torch.Generator().manual_seed(1)

    def batch(b, vol):
        return [torch.rand(b, 1, 200, generator=g).pin_memory(), torch.rand(b, 64, 64, 3, generator=g).pin_memory(),
                torch.rand((b,) + vol + (3,), generator=g).pin_memory()]

    calls = [batch(2, (4, 48, 48)) for _ in range(7)] + [batch(3, (3, 40, 56))] + [batch(2, (4, 48, 48)) for _ in range(3)]
    with torch.no_grad():
        want = [m([t.cuda() for t in xs]).clone() for xs in calls]
        torch.cuda.synchronize()
        m.keep_output_on_device = True
        got = [m(list(xs)) for xs in calls]          # all enqueued back to back, nothing read in between
        torch.cuda.synchronize()
        m.keep_output_on_device = False
    for i, (a, b) in enumerate(zip(got, want)):
        assert torch.equal(a, b), f"call {i}: host-staged forward differs from the device-input forward"
    assert len(m._stage_ring) == m.host_staging_depth


def test_forward_is_cuda_graph_capturable():
    """hn_forward is a pure stream-ordered launch sequence (no host sync, no allocation, programmatic dependent launches):
    a caller may capture it in a CUDA graph; the replay must reproduce the eager logits bit for bit (DESIGN.md 5c,
    tools/graph_forward.py)."""
    torch.manual_seed(0)
    kw = dict(n_modalities=3, channel_dims=[200, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, l_c=128, l_d=128, depth=2)
    m = HealNet(**kw).eval().cuda()
    xs = [torch.rand(2, 1, 200, device="cuda"), torch.rand(2, 64, 64, 3, device="cuda"),
          torch.rand(2, 4, 48, 48, 3, device="cuda")]
    with torch.no_grad():
        want = m(list(xs)).clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            m(list(xs))
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            got = m(list(xs))
        for t in xs:
            t.mul_(0.5)                    # new data in the captured input buffers
        want2 = None
        g.replay()
        torch.cuda.synchronize()
        replayed = got.clone()
        want2 = m(list(xs))
    assert not torch.equal(want, want2)
    assert torch.equal(replayed, want2), "graph replay differs from the eager forward on the same inputs"


def test_capture_graph_helper_matches_the_eager_forward():
    """HealNet.capture_graph(): replay on new host / device inputs equals the eager forward; shape changes are refused."""
    torch.manual_seed(0)
    kw = dict(n_modalities=3, channel_dims=[200, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, l_c=128, l_d=128, depth=2)
    m = HealNet(**kw).eval().cuda()
    g = torch.Generator().manual_seed(3)
    mk = lambda: [torch.rand(1, 1, 200, generator=g), None, torch.rand(1, 4, 48, 48, 3, generator=g)]
    with torch.no_grad():
        run = m.capture_graph([t if t is None else t.cuda() for t in mk()])
        for k in range(3):
            xs = mk()
            ins = xs if k % 2 else [t if t is None else t.cuda() for t in xs]       # host tensors and device tensors
            got = run(ins).clone()
            want = m([t if t is None else t.cuda() for t in xs])
            assert torch.equal(got, want)
        with pytest.raises(ValueError):
            run([torch.rand(2, 1, 200), None, torch.rand(2, 4, 48, 48, 3)])
