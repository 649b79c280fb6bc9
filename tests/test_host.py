"""CPU: host-side logic of the drop-in boundary — library export table, constructor / state_dict parity
with the reference, argument validation of the C ABI entry points that need no GPU."""
import ctypes
import os
import re

import pytest
import torch

import healnet_b200
from healnet_b200 import _lib
from healnet_b200 import HealNet, Attention
from conftest import ROOT

CASES = ["tri_small", "omic_wsi_tied", "plain_no_head", "two_ltiles", "prod_ucec", "wide_heads"]


def test_library_exports_every_declared_symbol():
    lib = healnet_b200.load_library()
    header = open(os.path.join(ROOT, "include", "healnet_b200.h")).read()
    declared = set(re.findall(r"HN_API\s+[\w\s\*]+?\b(hn_\w+)\s*\(", header))
    assert len(declared) >= 17
    assert declared == set(_lib.SIGNATURES), "ctypes table and header drifted apart"
    for name in declared:
        assert getattr(lib, name) is not None


@pytest.mark.parametrize("name", CASES)
def test_state_dict_layout_and_init_match_reference(golden, name):
    """Same keys, shapes AND same initial values for the same torch seed (parameter creation order of
    healnet.py:143-185), so reference checkpoints load both ways (explainer.py:359,400)."""
    meta, sd, _, _, _ = golden(name)
    torch.manual_seed(meta["seed"])
    model = HealNet(**meta["kwargs"])
    mine = model.state_dict()
    assert list(mine.keys()) == list(sd.keys())
    for k in sd:
        assert tuple(mine[k].shape) == tuple(sd[k].shape), k
        s, a = meta["init_sig"][k]
        assert abs(float(mine[k].double().sum()) - s) <= 1e-9 * max(1.0, abs(s)), k
        assert abs(float(mine[k].double().abs().sum()) - a) <= 1e-9 * max(1.0, a), k
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


def test_weight_tying_polarity():
    """healnet.py:161: layers >= 1 share modules when weight_tie_layers=True, layer 0 never does; tied
    cross-FF is also shared across modalities (single cache key)."""
    kw = dict(n_modalities=2, channel_dims=[4, 3], num_spatial_axes=[1, 2], out_dims=2, depth=3, l_c=8, l_d=16)
    m = HealNet(weight_tie_layers=True, **kw)
    assert m.layers[1][0] is m.layers[2][0] and m.layers[0][0] is not m.layers[1][0]
    assert m.layers[1][1] is m.layers[1][3] and m.layers[0][1] is not m.layers[0][3]
    assert m.layers[1][-1][0] is m.layers[2][-1][0]
    assert len(m.state_dict()) == len(HealNet(**kw).state_dict())
    u = HealNet(**kw)
    assert u.layers[1][0] is not u.layers[2][0]


def test_constructor_asserts():
    with pytest.raises(AssertionError):  # healnet.py:121-122, reference test_healnet.py:61-67
        HealNet(n_modalities=1, channel_dims=[2189, 100], num_spatial_axes=[1, 1], out_dims=4)
    with pytest.raises(AssertionError):
        HealNet(n_modalities=2, channel_dims=[5], num_spatial_axes=[1, 1], out_dims=4)
    with pytest.raises(TypeError):  # keyword-only, healnet.py:17
        HealNet(1, [5], [1], 4)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = HealNet(n_modalities=1, channel_dims=[5], num_spatial_axes=[1], out_dims=2, l_c=8, l_d=16)
    with pytest.raises(healnet_b200.HealNetLibraryError):
        m([torch.rand(2, 3, 5)])
    with pytest.raises(healnet_b200.HealNetLibraryError):
        Attention(16, 5)(torch.rand(2, 8, 16), context=torch.rand(2, 3, 5))


def test_forward_argument_checks_precede_device_use():
    m = HealNet(n_modalities=2, channel_dims=[5, 3], num_spatial_axes=[1, 2], out_dims=2, l_c=8, l_d=16)
    with pytest.raises((AssertionError, healnet_b200.HealNetLibraryError)):
        m([torch.rand(2, 3, 5), torch.rand(2, 4, 3)])  # modality 2 declared with 2 axes (healnet.py:207-208)


def _desc(**over):
    d = _lib.hn_desc()
    base = dict(n_modalities=2, depth=2, l_c=16, l_d=32, x_heads=2, cross_dim_head=16, l_heads=2,
                latent_dim_head=16, num_freq_bands=2, out_dims=3, self_per_cross_attn=1, snn=1,
                final_classifier_head=1, fourier_encode_data=1, max_freq=10.0)
    base.update(over)
    for k, v in base.items():
        setattr(d, k, v)
    d.channel_dims[0], d.channel_dims[1] = 7, 3
    d.num_spatial_axes[0], d.num_spatial_axes[1] = 1, 2
    return d


def test_hn_create_validation_and_workspace_sizing():
    lib = healnet_b200.load_library()
    h = ctypes.c_void_p()
    d = _desc()
    assert lib.hn_create(ctypes.byref(d), ctypes.byref(h)) == 0
    sizes = (ctypes.c_int * (_lib.HN_MAX_MODALITIES * _lib.HN_MAX_AXES))()
    sizes[0] = 1
    sizes[4], sizes[5] = 20, 30
    w1 = lib.hn_workspace_bytes(h, 1, sizes)
    w4 = lib.hn_workspace_bytes(h, 4, sizes)
    assert 0 < w1 < w4
    sizes[4] = 200   # 6000 tokens: past the short-axis (precise) regime, streaming small-context path
    w6k = lib.hn_workspace_bytes(h, 4, sizes)
    sizes[4] = 2000
    assert lib.hn_workspace_bytes(h, 4, sizes) > w6k > 0
    # no weights registered yet: packing must refuse, with a message
    assert lib.hn_pack_weights(h, None) != 0
    assert b"registered" in lib.hn_last_error()
    # wrong tensor count for a slot
    arr = (ctypes.c_void_p * 3)(1, 2, 3)
    assert lib.hn_set_weights(h, 0, 0, arr, 3) != 0
    assert lib.hn_set_weights(h, 5, 0, arr, 3) != 0
    assert lib.hn_destroy(h) == 0
    for bad in (dict(self_per_cross_attn=2), dict(cross_dim_head=129), dict(latent_dim_head=129), dict(depth=0),
                dict(n_modalities=17),
                dict(l_heads=0)):
        h2 = ctypes.c_void_p()
        d2 = _desc(**bad)
        assert lib.hn_create(ctypes.byref(d2), ctypes.byref(h2)) != 0, bad
        assert len(lib.hn_last_error()) > 0


def test_token_sharding_entry_points_validate_on_the_host():
    """hn_exchange_bytes / hn_workspace_bytes_split / hn_set_exchange are pure host logic: sizes scale as documented
    and bad arguments are refused with a message (no GPU needed)."""
    lib = healnet_b200.load_library()
    h = ctypes.c_void_p()
    d = _desc()
    assert lib.hn_create(ctypes.byref(d), ctypes.byref(h)) == 0
    b1, b4 = lib.hn_exchange_bytes(h, 1), lib.hn_exchange_bytes(h, 4)
    assert 256 < b1 < b4 and (b4 - 256) % 512 == 0            # header + two 256-byte-aligned slots
    assert lib.hn_exchange_bytes(h, 0) == 0
    sizes = (ctypes.c_int * (_lib.HN_MAX_MODALITIES * _lib.HN_MAX_AXES))()
    sizes[0] = 1
    sizes[4], sizes[5] = 200, 300                              # 60 000 tokens: streaming path, shardable
    full = lib.hn_workspace_bytes(h, 4, sizes)
    tc = (ctypes.c_long * _lib.HN_MAX_MODALITIES)()
    tc[1] = 60000 // 8
    part = lib.hn_workspace_bytes_split(h, 4, sizes, tc)
    assert 0 < part < full                                     # an eighth of the context rows, same latent side
    tc[1] = 60000                                              # "all tokens" = replicated: same plan as the full call
    assert abs(lib.hn_workspace_bytes_split(h, 4, sizes, tc) - full) <= 4096
    sizes[4], sizes[5] = 20, 30                                # 600 tokens: short axes are replicated, never sharded
    tc[1] = 100
    assert lib.hn_workspace_bytes_split(h, 4, sizes, tc) == 0
    assert b"replicated" in lib.hn_last_error()
    bufs = (ctypes.c_void_p * 9)(*[256 * (i + 1) for i in range(9)])
    assert lib.hn_set_exchange(h, 0, 9, bufs, b4) != 0         # one node: at most 8 ranks
    assert lib.hn_set_exchange(h, 3, 2, bufs, b4) != 0         # rank outside the world
    bufs[1] = 257
    assert lib.hn_set_exchange(h, 0, 2, bufs, b4) != 0         # unaligned peer buffer
    assert b"aligned" in lib.hn_last_error()
    assert lib.hn_set_exchange(h, 0, 1, None, 0) == 0          # world 1: sharding off
    assert lib.hn_destroy(h) == 0


def test_attention_split_heuristic_covers_the_chip():
    lib = healnet_b200.load_library()
    # cfg 1 volume modality: 4 samples x 8 heads x 4 latent tiles = 128 base CTAs -> must split the token axis
    ns = lib.hn_op_attention_nsplit(4, 512, 8, 602112, 0)
    assert ns >= 2 and 4 * 8 * 4 * ns >= 2 * 148
    assert lib.hn_op_attention_nsplit(4, 512, 8, 1, 0) == 1
    assert lib.hn_op_attention_nsplit(1, 25, 1, 64 * 16, 0) == 1
    # small-context kernel: one CTA per SM, 3 row blocks per CTA -> 4 * ceil(32 / 3) * ns CTAs, close to whole waves
    ns = lib.hn_op_attention_nsplit(4, 512, 8, 602112, 32)
    ctas = 4 * 11 * ns
    assert ns >= 3 and ctas >= 148 and (ctas % 148 == 0 or ctas % 148 >= 110)
