"""Drop-in `HealNet` / `Attention` modules backed by libhealnet_b200.so (hand-written sm_100a kernels).

Mirrors the reference's model surface (healnet/models/healnet.py): the keyword-only `HealNet(...)`
constructor (:15-38) with its asserts (:121-122), `forward(tensors, mask, return_embeddings, verbose)`
(:190-195), `get_attention_weights()` (:252-262), `Attention(query_dim, context_dim, heads, dim_head,
dropout)` (:370), and — because checkpoints move both ways (explainer.py:359,400) — the exact parameter
names / shapes / creation order of the reference's `state_dict()`. The modules below own parameters
only; all arithmetic of the forward pass happens in the CUDA library behind the C ABI
(include/healnet_b200.h). There is no PyTorch or CPU fallback: without the library or a CUDA device the
forward raises.
"""
from __future__ import annotations

import ctypes
import warnings
from typing import List, Optional, Sequence, Union

import torch
from torch import nn

from . import _lib
from .distributed import token_shard_bounds
from ._lib import HN_MAX_AXES, HN_MAX_MODALITIES, check, hn_desc, load_library


# ----------------------------------------------------------------------------------------------------------
# Parameter containers. Names of attributes are part of the checkpoint format.
# ----------------------------------------------------------------------------------------------------------
class _FusedOnly(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover - guard
        raise RuntimeError(f"{type(self).__name__} is evaluated inside the fused HealNet forward "
                           "(libhealnet_b200.so); call the owning HealNet module instead")


class _GateMarker(_FusedOnly):
    """Stands where the reference puts its SELU()/GELU() gate module (healnet.py:323-331) — no parameters;
    keeps `net.2` the index of the second Linear."""

    def __init__(self, kind: str):
        super().__init__()
        self.kind = kind

    def extra_repr(self) -> str:
        return f"a * {self.kind}(gates)"


class _MeanOverLatents(_FusedOnly):
    """Stands where the reference puts Reduce('b n d -> b d', 'mean') (healnet.py:182) — keeps LayerNorm /
    Linear at indices 1 / 2 of `to_logits`."""


class FeedForward(_FusedOnly):
    """Parameters of the gated feed-forward (healnet.py:339-351): Linear(D, 8D) -> a*act(g) -> Linear(4D, D)."""

    def __init__(self, dim: int, mult: int = 4, dropout: float = 0., snn: bool = False):
        super().__init__()
        self.net = nn.Sequential(
            nn.Linear(dim, dim * mult * 2),
            _GateMarker("selu" if snn else "gelu"),
            nn.Linear(dim * mult, dim),
            nn.Dropout(dropout),
        )


class Attention(nn.Module):
    """Multi-head attention with the reference's 0.5 softmax temperature and LeakyReLU output projection
    (healnet.py:369-426). Usable stand-alone: `Attention(query_dim, context_dim)(x, context=ctx, mask=m)`
    runs `hn_attention_forward`; inside HealNet it only holds the parameters."""

    def __init__(self, query_dim: int, context_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.):
        super().__init__()
        inner = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        self.query_dim = query_dim
        self.context_dim = context_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_kv = nn.Linear(context_dim, inner * 2, bias=False)
        self.dropout = nn.Dropout(dropout)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.LeakyReLU(negative_slope=1e-2))
        self.attn_weights = None  # the streaming kernel never materialises the attention matrix
        self._ws = None

    def forward(self, x: torch.Tensor, context: Optional[torch.Tensor] = None,
                mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        lib = load_library()
        dev = _compute_device(self.to_q.weight)
        if self.dropout.p > 0 and self.training:
            raise NotImplementedError("attention dropout > 0 in training mode is not supported (forward-only path)")
        ret_dev, ret_dtype = x.device, x.dtype
        xs = x.detach().to(device=dev, dtype=torch.float32).contiguous()
        b, n_q, qd = xs.shape
        if qd != self.query_dim:
            raise ValueError(f"x has width {qd}, expected query_dim={self.query_dim}")
        cs = None
        n_ctx, cd = n_q, qd
        if context is not None:
            cs = context.detach().to(device=dev, dtype=torch.float32).contiguous()
            if cs.dim() != 3 or cs.shape[0] != b or cs.shape[2] != self.context_dim:
                raise ValueError(f"context must be (batch, n, {self.context_dim}); got {tuple(cs.shape)}")
            n_ctx, cd = cs.shape[1], cs.shape[2]
        ms = None
        if mask is not None:
            ms = mask.detach().to(device=dev).reshape(b, -1).to(torch.uint8).contiguous()
            if ms.shape[1] != n_ctx:
                raise ValueError(f"mask has {ms.shape[1]} tokens per sample, context has {n_ctx}")
        w = [_f32_dev(p, dev) for p in (self.to_q.weight, self.to_kv.weight, self.to_out[0].weight,
                                        self.to_out[0].bias)]
        out = torch.empty(b, n_q, qd, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            need = lib.hn_attention_workspace_bytes(b, n_q, n_ctx, qd, cd, self.heads, self.dim_head)
            if need == 0:
                raise _lib.HealNetLibraryError("hn_attention_workspace_bytes rejected the shape")
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(need, device=dev, dtype=torch.uint8)
                self._ws_sig = None
            # the packed weights stay in the workspace: skip the packing launches while shapes and parameters are unchanged
            sig = (b, n_q, n_ctx, qd, cd, cs is None, self._ws.data_ptr(),
                   tuple((p.data_ptr(), p._version) for p in (self.to_q.weight, self.to_kv.weight, self.to_out[0].weight)))
            packed = 1 if getattr(self, "_ws_sig", None) == sig else 0
            stream = torch.cuda.current_stream(dev).cuda_stream
            check(lib.hn_attention_forward_cached(b, n_q, n_ctx, qd, cd, self.heads, self.dim_head, xs.data_ptr(),
                                                  cs.data_ptr() if cs is not None else None, w[0].data_ptr(),
                                                  w[1].data_ptr(), w[2].data_ptr(), w[3].data_ptr(),
                                                  ms.data_ptr() if ms is not None else None, out.data_ptr(),
                                                  self._ws.data_ptr(), self._ws.numel(), packed, stream),
                  "hn_attention_forward")
            self._ws_sig = sig
        return out.to(device=ret_dev, dtype=ret_dtype)


class PreNorm(_FusedOnly):
    """LayerNorm(s) in front of `fn` (healnet.py:306-321); parameters only."""

    def __init__(self, dim: int, fn: nn.Module, context_dim: Optional[int] = None):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)
        self.norm_context = nn.LayerNorm(context_dim) if context_dim is not None else None


def _memoize(factory):
    """Constructor-time weight tying: with `_cache=True` the same module instance is returned for a key
    (behaviour of the reference's cache_fn, healnet.py:278-290)."""
    made = {}

    def get(_cache: bool = True, key=None):
        if not _cache:
            return factory()
        if key not in made:
            made[key] = factory()
        return made[key]

    return get


def _compute_device(param: torch.Tensor) -> torch.device:
    if param.device.type == "cuda":
        return param.device
    if not torch.cuda.is_available():
        raise _lib.HealNetLibraryError(
            "healnet_b200 needs a CUDA device (NVIDIA B200, sm_100a): there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _f32_dev(p: torch.Tensor, dev: torch.device) -> torch.Tensor:
    t = p.detach()
    if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(device=dev, dtype=torch.float32).contiguous()
    return t


# ----------------------------------------------------------------------------------------------------------
class _FusedForward(torch.autograd.Function):
    """autograd node of the fused forward: forward = hn_forward_train (records a tape), backward = hn_backward
    (reference training step: healnet/main.py:426-467 calls loss.backward() through HealNet.forward). The parameters
    are passed as inputs so that autograd routes their gradients; the data tensors get none (the reference never
    asks for them)."""

    @staticmethod
    def forward(ctx, model, call, *params):
        out, state = model._launch_train(call, params)
        ctx.model, ctx.state = model, state
        return out

    @staticmethod
    def backward(ctx, grad_out):
        grads = ctx.model._backward(ctx.state, grad_out)
        ctx.state = None
        return (None, None, *grads)


# ----------------------------------------------------------------------------------------------------------
class HealNet(nn.Module):
    """HEALNet fusion model — same constructor, parameters and forward contract as the reference
    (healnet/models/healnet.py:14-262); the forward pass runs as sm_100a kernels through the C ABI.

    Differences, all deliberate and documented in DESIGN.md:
      * forward does not mutate the caller's list (the reference does, :222);
      * the attention matrices are never materialised, so `get_attention_weights()` returns `None`
        entries (the reference keeps b*h*L*N floats per Attention module, :420);
      * with gradients enabled the forward records a tape and `loss.backward()` runs the library's backward
        (hn_backward): parameter gradients only, fp32 CUDA parameters, dropout 0.
    Quirks that ARE reproduced: softmax temperature 0.5 on top of dim_head**-0.5 (:375,:419); the latent
    self-attention block runs after every modality (:228-245); a missing / mismatching modality skips only
    its cross-attention + cross-FF unless `verbose=True` (:229-239); layer-0 weights are never tied (:161).
    """

    def __init__(
        self,
        *,
        n_modalities: int,
        channel_dims: List,
        num_spatial_axes: List,
        out_dims: int,
        depth: int = 3,
        num_freq_bands: int = 2,
        max_freq: float = 10.,
        l_c: int = 128,
        l_d: int = 128,
        x_heads: int = 8,
        l_heads: int = 8,
        cross_dim_head: int = 64,
        latent_dim_head: int = 64,
        attn_dropout: float = 0.,
        ff_dropout: float = 0.,
        weight_tie_layers: bool = False,
        fourier_encode_data: bool = True,
        self_per_cross_attn: int = 1,
        final_classifier_head: bool = True,
        snn: bool = True,
    ):
        super().__init__()
        assert len(channel_dims) == len(num_spatial_axes), 'input channels and input axis must be of the same length'
        assert len(num_spatial_axes) == n_modalities, 'input axis must be of the same length as the number of modalities'

        self.input_axes = list(num_spatial_axes)
        self.input_channels = list(channel_dims)
        self.max_freq = max_freq
        self.num_freq_bands = num_freq_bands
        self.modalities = n_modalities
        self.self_per_cross_attn = self_per_cross_attn
        self.fourier_encode_data = fourier_encode_data
        self._hparams = dict(depth=depth, l_c=l_c, l_d=l_d, x_heads=x_heads, l_heads=l_heads,
                             cross_dim_head=cross_dim_head, latent_dim_head=latent_dim_head, out_dims=out_dims,
                             snn=snn, final_classifier_head=final_classifier_head, attn_dropout=attn_dropout,
                             ff_dropout=ff_dropout)

        per_axis = (2 * num_freq_bands + 1) if fourier_encode_data else 0
        input_dims = [c + a * per_axis for c, a in zip(channel_dims, num_spatial_axes)]

        # Parameter creation order follows the reference (:143-185) so that a given torch seed yields the
        # same initial weights: latents; per layer [latent attn, latent ff, then per modality cross attn,
        # cross ff]; head last.
        self.latents = nn.Parameter(torch.randn(l_c, l_d))

        def cross_attn_factory(m):
            return lambda: PreNorm(l_d, Attention(l_d, input_dims[m], heads=x_heads, dim_head=cross_dim_head,
                                                  dropout=attn_dropout), context_dim=input_dims[m])

        get_cross_attn = [_memoize(cross_attn_factory(m)) for m in range(n_modalities)]
        get_latent_attn = _memoize(lambda: PreNorm(l_d, Attention(l_d, heads=l_heads, dim_head=latent_dim_head,
                                                                  dropout=attn_dropout)))
        get_cross_ff = _memoize(lambda: PreNorm(l_d, FeedForward(l_d, dropout=ff_dropout, snn=snn)))
        get_latent_ff = _memoize(lambda: PreNorm(l_d, FeedForward(l_d, dropout=ff_dropout, snn=snn)))

        self.layers = nn.ModuleList([])
        for i in range(depth):
            tie = i > 0 and weight_tie_layers
            latent_block = nn.ModuleList([])
            for block in range(self_per_cross_attn):
                latent_block.append(get_latent_attn(_cache=tie, key=block))
                latent_block.append(get_latent_ff(_cache=tie, key=block))
            entries = []
            for m in range(n_modalities):
                entries.append(get_cross_attn[m](_cache=tie))
                entries.append(get_cross_ff(_cache=tie))
            self.layers.append(nn.ModuleList([*entries, latent_block]))

        self.to_logits = nn.Sequential(
            _MeanOverLatents(),
            nn.LayerNorm(l_d),
            nn.Linear(l_d, out_dims),
        ) if final_classifier_head else nn.Identity()

        # native state (not part of the checkpoint)
        self._handle = None
        self._handle_dev = None
        self._weights_sig = None
        self._staged = []       # keeps staged fp32 device copies alive while the handle borrows them
        self._workspace = None
        self._copy_stream = None
        # inference calls with pinned host inputs: persistent device staging sets used round-robin. The copies of call
        # i wait only for the forward that last READ set i % depth (an event per set), not for all earlier compute, so
        # step i+1's host-to-device transfer runs under step i's kernels whatever the PCIe rate of the day is.
        self.host_staging_depth = 3
        self._stage_ring = None     # [depth] dicts: modality index -> device tensor
        self._stage_done = None     # [depth] events: the forward that consumed the set has been enqueued / finished
        self._stage_k = 0
        self._export_registered = False
        # token-axis sharding across GPUs (enable_token_sharding): (rank, world, min_tokens) or None
        self._token_shard = None
        self._exchange = None   # (own device ptr, [peer ptrs], bytes, max batch, handle)
        self._exchange_group = None
        self._xerr = None       # (pinned int32 flag, event): asynchronous read-back of the peer-wait time-out flag
        self.token_sharding_timeout_s = 30.0
        # opt-in: materialise every Attention module's softmax matrix on each forward (reference: always on,
        # healnet.py:420). Off by default: (b*h, L, N) fp32 is 9.87 GB per sample and layer at the README shapes.
        self.export_attention_weights = False
        self.export_attention_max_bytes = 16 << 30
        # serving loops: leave the result on the compute device even when the inputs came from the host, so the caller
        # can read it back asynchronously (pinned buffer + event) and enqueue the next batch meanwhile
        self.keep_output_on_device = False
        self.last_launch_count = 0
        self._warned = set()
        self._train_seq = 0      # training-mode forwards so far: only the latest one can be back-propagated
        # "tensor": heavy backward contractions on tensor cores (default); "fp32": the exact SIMT kernels they are
        # checked against (hn_set_backward_variant)
        self.backward_variant = "tensor"

    # ------------------------------------------------------------------------------------------------ native
    _NATIVE_DEFAULTS = dict(_handle=None, _handle_dev=None, _weights_sig=None, _staged=(), _workspace=None,
                            _copy_stream=None, _export_registered=False, _exchange=None, _token_shard=None,
                            _exchange_group=None, _xerr=None, _stage_ring=None, _stage_done=None, _stage_k=0)

    def __getstate__(self):
        """copy.deepcopy / pickle / torch.save(model): the native handle, staged weights and workspace are
        per-object caches and are rebuilt lazily by the copy."""
        state = dict(self.__dict__)
        for k, v in self._NATIVE_DEFAULTS.items():
            state[k] = [] if k == "_staged" else v
        state["_warned"] = set()
        return state

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _release_exchange(self):
        if getattr(self, "_exchange", None) is not None:
            lib = load_library()
            mine, bufs = self._exchange[0], self._exchange[1]
            if torch.cuda.is_available():
                torch.cuda.synchronize()   # no kernel of ours may still be reading a peer mapping
            for p in bufs:
                if p != mine:
                    lib.hn_exchange_close(p)
            lib.hn_exchange_free(mine)
            self._exchange = None
            self._xerr = None

    def _release(self):
        self._release_exchange()
        if getattr(self, "_handle", None) is not None:
            load_library().hn_destroy(self._handle)
            self._handle = None

    def _desc(self) -> hn_desc:
        hp = self._hparams
        if self.modalities > HN_MAX_MODALITIES:
            raise ValueError(f"at most {HN_MAX_MODALITIES} modalities are supported")
        d = hn_desc()
        d.n_modalities = self.modalities
        d.depth = hp["depth"]
        d.l_c, d.l_d = hp["l_c"], hp["l_d"]
        d.x_heads, d.cross_dim_head = hp["x_heads"], hp["cross_dim_head"]
        d.l_heads, d.latent_dim_head = hp["l_heads"], hp["latent_dim_head"]
        d.num_freq_bands = self.num_freq_bands
        d.out_dims = hp["out_dims"]
        d.self_per_cross_attn = self.self_per_cross_attn
        d.snn = 1 if hp["snn"] else 0
        d.final_classifier_head = 1 if hp["final_classifier_head"] else 0
        d.fourier_encode_data = 1 if self.fourier_encode_data else 0
        d.max_freq = float(self.max_freq)
        for m in range(self.modalities):
            d.channel_dims[m] = int(self.input_channels[m])
            d.num_spatial_axes[m] = int(self.input_axes[m])
        return d

    def _slot_params(self):
        """[(layer, slot, [parameters in the order hn_set_weights documents])]"""
        M = self.modalities
        out = [(-1, 0, [self.latents])]
        if self._hparams["final_classifier_head"]:
            ln, lin = self.to_logits[1], self.to_logits[2]
            out.append((-1, 1, [ln.weight, ln.bias, lin.weight, lin.bias]))
        for l, layer in enumerate(self.layers):
            for m in range(M):
                pa, pf = layer[2 * m], layer[2 * m + 1]
                out.append((l, 2 * m, [pa.norm.weight, pa.norm.bias, pa.norm_context.weight, pa.norm_context.bias,
                                       pa.fn.to_q.weight, pa.fn.to_kv.weight, pa.fn.to_out[0].weight,
                                       pa.fn.to_out[0].bias]))
                out.append((l, 2 * m + 1, [pf.norm.weight, pf.norm.bias, pf.fn.net[0].weight, pf.fn.net[0].bias,
                                           pf.fn.net[2].weight, pf.fn.net[2].bias]))
            if self.self_per_cross_attn > 0:
                pa, pf = layer[-1][0], layer[-1][1]
                out.append((l, 2 * M, [pa.norm.weight, pa.norm.bias, pa.fn.to_q.weight, pa.fn.to_kv.weight,
                                       pa.fn.to_out[0].weight, pa.fn.to_out[0].bias]))
                out.append((l, 2 * M + 1, [pf.norm.weight, pf.norm.bias, pf.fn.net[0].weight, pf.fn.net[0].bias,
                                           pf.fn.net[2].weight, pf.fn.net[2].bias]))
        return out

    def _sync_native(self, dev: torch.device, stream: int):
        """Creates the native handle on first use and re-registers / repacks weights whenever a parameter's
        storage, device or in-place version changed (optimizer steps, load_state_dict, .to())."""
        lib = load_library()
        if self.self_per_cross_attn not in (0, 1):
            # the reference unpacks `self_attn, self_ff = layer[-1]` (healnet.py:242)
            raise ValueError("too many values to unpack: self_per_cross_attn must be 0 or 1")
        if self._handle is not None and self._handle_dev != dev:
            self._release()
        if self._handle is None:
            hp = ctypes.c_void_p()
            d = self._desc()
            check(lib.hn_create(ctypes.byref(d), ctypes.byref(hp)), "hn_create")
            self._handle, self._handle_dev, self._weights_sig = hp, dev, None
        slots = self._slot_params()
        sig = tuple((p.data_ptr(), p._version, p.dtype, p.device) for _, _, ps in slots for p in ps)
        if sig == self._weights_sig:
            return
        staged = []
        for layer, slot, ps in slots:
            ts = [_f32_dev(p, dev) for p in ps]
            staged.extend(ts)
            arr = (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
            check(lib.hn_set_weights(self._handle, layer, slot, arr, len(ts)), "hn_set_weights")
        self._staged = staged
        check(lib.hn_pack_weights(self._handle, stream), "hn_pack_weights")
        self._weights_sig = sig

    def _warn_once(self, key, msg):
        if key not in self._warned:
            self._warned.add(key)
            warnings.warn(msg, RuntimeWarning, stacklevel=3)

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self,
                tensors: List[Union[torch.Tensor, None]],
                mask: Optional[torch.Tensor] = None,
                return_embeddings: bool = False,
                verbose: bool = False):
        lib = load_library()
        hp = self._hparams
        if self.training and (hp["attn_dropout"] > 0 or hp["ff_dropout"] > 0):
            raise NotImplementedError("dropout > 0 in training mode is not supported (forward-only path)")
        dev = _compute_device(self.latents)
        M = self.modalities
        n_given = len(tensors)
        missing_idx = [i for i, t in enumerate(tensors) if t is None]
        if verbose:
            print(f"Missing modalities indices: {missing_idx}")

        batch = None
        ret_dev, ret_dtype = None, None
        staged: List[Optional[torch.Tensor]] = [None] * M
        ready: List[Optional[torch.cuda.Event]] = [None] * M   # per-modality "copy finished" events (host inputs)
        axis_sizes = (ctypes.c_int * (HN_MAX_MODALITIES * HN_MAX_AXES))()
        axis_tokens = [0] * M   # full token count of every given modality
        tok_begin = tok_count = None
        if self._token_shard is not None:
            tok_begin, tok_count = [0] * M, [0] * M
        # inputs that all arrive in one 16-bit floating type are consumed as they are (hn_set_io_dtype): no widened
        # copy, half the host-to-device bytes; anything else is staged as fp32 (the reference's dtype follows its
        # inputs, healnet.py:212)
        # (training-mode calls keep freshly allocated copies: the backward pass may still need them)
        ring_k = None
        if not torch.is_grad_enabled() and self.host_staging_depth >= 2:
            if self._stage_ring is None or len(self._stage_ring) != self.host_staging_depth:
                self._stage_ring = [dict() for _ in range(self.host_staging_depth)]
                self._stage_done = [None] * self.host_staging_depth
                self._stage_k = 0
            ring_k = self._stage_k
            self._stage_k = (ring_k + 1) % self.host_staging_depth
            self._stage_waited = False
        given = {t.dtype for t in tensors[:M] if t is not None}
        io_dtype = given.pop() if len(given) == 1 and next(iter(given)) in (torch.bfloat16, torch.float16) else torch.float32
        for i in range(min(n_given, M)):
            data = tensors[i]
            if data is None:
                continue
            b, *axis, c = data.shape
            assert len(axis) == self.input_axes[i], (f'input data for modality {i + 1} must hav'
                                                     f' the same number of axis as the input axis parameter')
            batch = b  # the reference takes the batch size from the last modality it encodes (:206,:225)
            ret_dev, ret_dtype = data.device, data.dtype
            if c != self.input_channels[i]:
                # the reference fails inside its try/except and silently skips this modality (:235-239)
                self._warn_once(("chan", i), f"modality {i}: got {c} channels, model expects "
                                f"{self.input_channels[i]}; its cross-attention is skipped (reference behaviour)")
                continue
            if len(axis) > HN_MAX_AXES:
                raise ValueError(f"at most {HN_MAX_AXES} spatial axes per modality are supported")
            axis_tokens[i] = 1
            for a, s in enumerate(axis):
                axis_sizes[i * HN_MAX_AXES + a] = int(s)
                axis_tokens[i] *= int(s)
            data = data.detach()
            if self._token_shard is not None and axis_tokens[i] >= self._token_shard[2]:
                # token-axis sharding: stage only this rank's slice of a long modality (positions stay global)
                rank, world, _ = self._token_shard
                lo, hi = token_shard_bounds(axis_tokens[i], world, rank)
                tok_begin[i], tok_count[i] = lo, hi - lo
                data = data.reshape(b, axis_tokens[i], c)[:, lo:hi]
            staged[i], ready[i] = self._stage_input(data, dev, io_dtype, None if ring_k is None else (ring_k, i))
        if batch is None:
            # reference: `b` is unbound -> UnboundLocalError at :225
            raise UnboundLocalError("cannot infer the batch size: every modality is missing")
        for i, t in enumerate(staged):
            if t is not None and t.shape[0] != batch:
                raise ValueError("all modalities must share the batch dimension")
        # `verbose=True` turns a missing modality into a full skip (self-attention included, :229-232)
        skip_self = [False] * M
        if verbose:
            for i in missing_idx:
                if i < M:
                    print(f"Skipping update in fusion layer for missing modality {i + 1}")
                    skip_self[i] = True

        mask_dev, mask_tokens = None, 0
        if mask is not None:
            mask_dev = mask.detach().to(device=dev).reshape(mask.shape[0], -1).to(torch.uint8).contiguous()
            if mask_dev.shape[0] != batch:
                raise ValueError("mask must have the batch dimension of the inputs")
            mask_tokens = int(mask_dev.shape[1])
            if tok_count is not None:
                # a mask addresses one token axis: cut it like the (single) sharded modality of that length
                cuts = {(tok_begin[i], tok_count[i]) for i, t in enumerate(staged)
                        if t is not None and tok_count[i] > 0 and axis_tokens[i] == mask_tokens}
                if len(cuts) == 1:
                    lo, cnt = cuts.pop()
                    mask_dev = mask_dev[:, lo:lo + cnt].contiguous()
                    mask_tokens = cnt
            for i, t in enumerate(staged):
                if t is not None:
                    n_tok = t.numel() // (batch * t.shape[-1])
                    if n_tok != mask_tokens and mask_tokens != 1:
                        # the reference's masked_fill_ raises on the shape mismatch and the modality is skipped
                        self._warn_once(("mask", i), f"modality {i}: mask has {mask_tokens} tokens, modality has "
                                        f"{n_tok}; its cross-attention is skipped (reference behaviour)")
                        staged[i] = None
        nan_rows = None
        if mask_dev is not None and mask_tokens == 1:
            # a (b, 1) mask broadcasts over the token axis of EVERY modality (healnet.py:411-415): True keeps all tokens
            # (a no-op), False masks all of them, and a fully masked softmax row is NaN in the reference
            # (-finfo.max / 0.5 = -inf), which then spreads over that sample's whole latent array. Same result here:
            # run unmasked, NaN the samples whose mask is False.
            nan_rows = mask_dev[:, 0] == 0
            mask_dev, mask_tokens = None, 0

        want_latents = return_embeddings or not hp["final_classifier_head"]
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            self._sync_native(dev, stream)
            check(lib.hn_set_io_dtype(self._handle, {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[io_dtype]),
                  "hn_set_io_dtype")
            params = [p for p in self.parameters()]
            differentiable = torch.is_grad_enabled() and any(p.requires_grad for p in params)
            if differentiable:
                # the library's backward covers fp32 parameters resident on the compute device; anything else runs the
                # plain forward and says (once) that its output carries no autograd graph
                why = None
                if any(p.device != dev or p.dtype != torch.float32 for p in params):
                    why = "parameters are not fp32 tensors on the CUDA device"
                elif tok_count is not None and any(c > 0 for c in tok_count):
                    why = "the token axis is sharded across GPUs"
                elif self.export_attention_weights:
                    why = "export_attention_weights is on"
                if why is not None:
                    self._warn_once(("nograd", why), f"HealNet forward under autograd, but {why}: the output carries "
                                    "no autograd graph (use torch.no_grad() for inference to silence this)")
                    differentiable = False
            if differentiable:
                call = dict(lib=lib, staged=staged, ready=ready, axis_sizes=axis_sizes, skip_self=skip_self,
                            mask_dev=mask_dev, mask_tokens=mask_tokens, batch=batch, want_latents=want_latents, dev=dev)
                out = _FusedForward.apply(self, call, *params)
            else:
                out = self._launch(lib, staged, ready, axis_sizes, skip_self, mask_dev, mask_tokens, batch,
                                   want_latents, dev, stream, tok_begin, tok_count)
            if ring_k is not None and any(e is not None for e in ready):
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(dev))
                self._stage_done[ring_k] = done
            if nan_rows is not None and any(t is not None for t in staged):
                out = torch.where(nan_rows.view(-1, *([1] * (out.dim() - 1))), torch.full_like(out, float("nan")), out)
        if ret_dtype is not None and not ret_dtype.is_floating_point:
            ret_dtype = torch.float32
        if self.keep_output_on_device:
            return out.to(dtype=ret_dtype)
        return out.to(device=ret_dev, dtype=ret_dtype)

    def _stage_input(self, data: torch.Tensor, dev: torch.device, dtype: torch.dtype = torch.float32, slot=None):
        """-> (contiguous device tensor of `dtype`, event or None). Pinned host tensors are copied on a side stream and
        the forward waits for each modality's copy only where it first reads it (hn_forward_ex), so the transfer
        of a large late modality overlaps the work on the earlier ones. `slot` = (staging set, modality): copy into
        the persistent buffer of that set (see host_staging_depth) instead of a fresh allocation."""
        if data.device == dev:
            return data.to(dtype=dtype).contiguous(), None
        # (a token-sharded slice of a pinned tensor is contiguous per sample: copied sample by sample, still async)
        per_sample = data.dim() == 3 and not data.is_contiguous() and all(data[i].is_contiguous() for i in range(data.shape[0]))
        if data.device.type == "cpu" and data.is_pinned() and data.dtype == dtype and (data.is_contiguous() or per_sample):
            if self._copy_stream is None or self._copy_stream.device != dev:
                self._copy_stream = torch.cuda.Stream(device=dev)
            cur = torch.cuda.current_stream(dev)
            if slot is not None:
                k, i = slot
                out = self._stage_ring[k].get(i)
                if out is None or out.shape != data.shape or out.dtype != dtype or out.device != dev:
                    out = torch.empty(data.shape, dtype=dtype, device=dev)
                    out.record_stream(self._copy_stream)
                    self._stage_ring[k][i] = out
                    self._copy_stream.wait_stream(cur)      # new memory: its earlier users on this stream must be done
                elif not self._stage_waited and self._stage_done[k] is not None:
                    self._copy_stream.wait_event(self._stage_done[k])   # the forward that last read this set
                self._stage_waited = True
            else:
                out = torch.empty(data.shape, dtype=dtype, device=dev)   # allocated on the compute stream
                self._copy_stream.wait_stream(cur)                          # ...whose earlier users must be done
            with torch.cuda.stream(self._copy_stream):
                if per_sample:
                    for i in range(data.shape[0]):
                        out[i].copy_(data[i], non_blocking=True)
                else:
                    out.copy_(data, non_blocking=True)
                if slot is None:
                    out.record_stream(self._copy_stream)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            return out, ev
        return data.to(device=dev, dtype=dtype, non_blocking=True).contiguous(), None

    def _launch(self, lib, staged, ready, axis_sizes, skip_self, mask_dev, mask_tokens, batch, want_latents, dev,
                stream, tok_begin=None, tok_count=None):
        hp = self._hparams
        M = self.modalities
        ptrs = (ctypes.c_void_p * HN_MAX_MODALITIES)()
        sizes = (ctypes.c_int * (HN_MAX_MODALITIES * HN_MAX_AXES))(*axis_sizes)
        for i in range(M):
            ptrs[i] = staged[i].data_ptr() if staged[i] is not None else None
            if staged[i] is None:
                for a in range(HN_MAX_AXES):   # sizing needs sane extents for absent modalities too
                    sizes[i * HN_MAX_AXES + a] = max(1, sizes[i * HN_MAX_AXES + a])
        split = tok_count is not None and any(c > 0 for c in tok_count)
        if split:
            tb = (ctypes.c_long * HN_MAX_MODALITIES)(*tok_begin)
            tc = (ctypes.c_long * HN_MAX_MODALITIES)(*tok_count)
            self._ensure_exchange(lib, batch)
            need = lib.hn_workspace_bytes_split(self._handle, batch, sizes, tc)
        else:
            need = lib.hn_workspace_bytes(self._handle, batch, sizes)
        if need == 0:
            raise _lib.HealNetLibraryError(f"hn_workspace_bytes failed: {_lib.last_error()}")
        if self._workspace is None or self._workspace.device != dev or self._workspace.numel() < need:
            self._workspace = None
            self._workspace = torch.empty(need, device=dev, dtype=torch.uint8)
        if want_latents:
            out = torch.empty(batch, hp["l_c"], hp["l_d"], device=dev, dtype=torch.float32)
            lat_ptr, log_ptr = out.data_ptr(), None
        else:
            out = torch.empty(batch, hp["out_dims"], device=dev, dtype=torch.float32)
            lat_ptr, log_ptr = None, out.data_ptr()
        skip = (ctypes.c_int * HN_MAX_MODALITIES)(*[1 if f else 0 for f in skip_self]) if any(skip_self) else None
        exported = self._register_attention_export(lib, staged, skip_self, batch, dev)
        events = None
        if any(e is not None for e in ready):
            events = (ctypes.c_void_p * HN_MAX_MODALITIES)()
            for i in range(M):
                events[i] = ready[i].cuda_event if (ready[i] is not None and staged[i] is not None) else None
        if split:
            self._poll_token_sharding(lib, None)      # an earlier forward's flag, if its read-back has landed
            check(lib.hn_forward_split(self._handle, batch, ptrs, events, sizes, tb, tc, skip,
                                       mask_dev.data_ptr() if mask_dev is not None else None,
                                       mask_tokens, lat_ptr, log_ptr, self._workspace.data_ptr(),
                                       self._workspace.numel(), stream), "hn_forward_split")
            self._poll_token_sharding(lib, stream)    # enqueue the read-back of this forward's flag
        else:
            check(lib.hn_forward_ex(self._handle, batch, ptrs, events, sizes, skip,
                                    mask_dev.data_ptr() if mask_dev is not None else None,
                                    mask_tokens, lat_ptr, log_ptr, self._workspace.data_ptr(), self._workspace.numel(),
                                    stream), "hn_forward")
        self.last_launch_count = lib.hn_last_launch_count(self._handle)
        for module, tensor in exported:   # later calls of a tied / repeated module win, as in the reference
            module.attn_weights = tensor
        return out

    # ------------------------------------------------------------------------------ training step (row f2)
    def _launch_train(self, call, params):
        """hn_forward_train on a workspace + tape owned by this call (they must outlive later forwards until the
        backward runs). Returns (output, state for _backward)."""
        lib, dev, batch = call["lib"], call["dev"], call["batch"]
        hp, M = self._hparams, self.modalities
        staged, ready = call["staged"], call["ready"]
        ptrs = (ctypes.c_void_p * HN_MAX_MODALITIES)()
        sizes = (ctypes.c_int * (HN_MAX_MODALITIES * HN_MAX_AXES))(*call["axis_sizes"])
        for i in range(M):
            ptrs[i] = staged[i].data_ptr() if staged[i] is not None else None
            if staged[i] is None:
                for a in range(HN_MAX_AXES):
                    sizes[i * HN_MAX_AXES + a] = max(1, sizes[i * HN_MAX_AXES + a])
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            need = lib.hn_workspace_bytes(self._handle, batch, sizes)
            tape_need = lib.hn_tape_bytes(self._handle, batch, sizes)
            if need == 0 or tape_need == 0:
                raise _lib.HealNetLibraryError(f"sizing the training forward failed: {_lib.last_error()}")
            workspace = torch.empty(need, device=dev, dtype=torch.uint8)
            tape = torch.empty(tape_need, device=dev, dtype=torch.uint8)
            if call["want_latents"]:
                out = torch.empty(batch, hp["l_c"], hp["l_d"], device=dev, dtype=torch.float32)
                lat_ptr, log_ptr = out.data_ptr(), None
            else:
                out = torch.empty(batch, hp["out_dims"], device=dev, dtype=torch.float32)
                lat_ptr, log_ptr = None, out.data_ptr()
            skip_self = call["skip_self"]
            skip = (ctypes.c_int * HN_MAX_MODALITIES)(*[1 if f else 0 for f in skip_self]) if any(skip_self) else None
            events = None
            if any(e is not None for e in ready):
                events = (ctypes.c_void_p * HN_MAX_MODALITIES)()
                for i in range(M):
                    events[i] = ready[i].cuda_event if (ready[i] is not None and staged[i] is not None) else None
            mask_dev = call["mask_dev"]
            check(lib.hn_forward_train(self._handle, batch, ptrs, events, sizes, skip,
                                       mask_dev.data_ptr() if mask_dev is not None else None, call["mask_tokens"],
                                       lat_ptr, log_ptr, workspace.data_ptr(), workspace.numel(), tape.data_ptr(),
                                       tape.numel(), stream), "hn_forward_train")
        self.last_launch_count = lib.hn_last_launch_count(self._handle)
        self._train_seq += 1
        state = dict(seq=self._train_seq, workspace=workspace, tape=tape, sizes=sizes, batch=batch, dev=dev,
                     want_latents=call["want_latents"], params=params, keep=(staged, mask_dev))
        return out, state

    def _backward(self, state, grad_out):
        """hn_backward: gradients of every parameter (zeros for parameters the forward did not use), in the order of
        `state['params']`."""
        lib = load_library()
        if state is None or state["seq"] != self._train_seq:
            raise RuntimeError("only the most recent training-mode forward of a HealNet module can be back-propagated "
                               "(one tape per native handle); call backward() before the next forward()")
        dev, batch, params = state["dev"], state["batch"], state["params"]
        # one zero-filled buffer for all gradients (one launch instead of one per parameter), 64-byte aligned views
        offs, total = {}, 0
        for p in params:
            if id(p) not in offs:
                offs[id(p)] = total
                total += (p.numel() + 15) // 16 * 16
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        grads = {id(p): flat[offs[id(p)]:offs[id(p)] + p.numel()].view(p.shape) for p in params}
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            for layer, slot, ps in self._slot_params():
                arr = (ctypes.c_void_p * len(ps))(*[grads[id(p)].data_ptr() for p in ps])
                check(lib.hn_set_grads(self._handle, layer, slot, arr, len(ps)), "hn_set_grads")
            check(lib.hn_set_backward_variant(self._handle, 1 if self.backward_variant == "fp32" else 0),
                  "hn_set_backward_variant")
            need = lib.hn_backward_scratch_bytes(self._handle, batch, state["sizes"])
            if need == 0:
                raise _lib.HealNetLibraryError(f"hn_backward_scratch_bytes failed: {_lib.last_error()}")
            scratch = torch.empty(need, device=dev, dtype=torch.uint8)
            g = grad_out.detach().to(device=dev, dtype=torch.float32).contiguous()
            ws, tape = state["workspace"], state["tape"]
            check(lib.hn_backward(self._handle, g.data_ptr() if state["want_latents"] else None,
                                  None if state["want_latents"] else g.data_ptr(), ws.data_ptr(), ws.numel(),
                                  tape.data_ptr(), tape.numel(), scratch.data_ptr(), scratch.numel(), stream),
                  "hn_backward")
            scratch.record_stream(torch.cuda.current_stream(dev))
        return [grads[id(p)] if p.requires_grad else None for p in params]

    # ------------------------------------------------------------------------------ token-axis sharding (row f4)
    def enable_token_sharding(self, group=None, min_tokens: int = 8192, max_batch: int = 8) -> None:
        """Shards the token axis of every modality with at least `min_tokens` tokens across the ranks of `group`
        (one process per GPU of ONE node): each rank streams its slice and the per-row softmax partials are merged
        over peer memory (hn_forward_split). Every rank must then call forward() with the same full inputs and gets
        the same (bit-identical) result. For batches too small to fill the GPUs by batch sharding."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            self._token_shard = None
            return
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > 8:
            raise ValueError("token sharding covers the GPUs of one node (<= 8 ranks)")
        # axes of up to 2048 tokens take the precise, replicated path in the library: never shard those
        self._token_shard = (rank, world, max(int(min_tokens), 2049))
        self._exchange_group, self._exchange_max_batch = group, int(max_batch)

    def disable_token_sharding(self) -> None:
        """Back to the replicated forward (the exchange buffers stay mapped until the module is released)."""
        self._token_shard = None

    def _poll_token_sharding(self, lib, stream, wait: bool = False) -> None:
        """Raises if a combine kernel of an earlier token-sharded forward gave up waiting for a peer (its outputs were
        overwritten with NaN). The flag travels back through a stream-ordered copy into pinned memory; `wait=False`
        only looks at copies that have already completed, so the hot path never synchronises."""
        if self._xerr is not None:
            flag, ev = self._xerr
            if wait:
                ev.synchronize()
            if ev.query():
                self._xerr = None
                if int(flag.item()) != 0:
                    raise _lib.HealNetLibraryError(
                        "token-sharded forward: a peer did not publish its attention partials within "
                        f"{self.token_sharding_timeout_s:.0f} s (token_sharding_timeout_s); the outputs of that "
                        "forward are NaN. All ranks must call forward() with the same inputs in the same order.")
        if stream is not None and self._xerr is None and self._exchange is not None:
            flag = torch.zeros(1, dtype=torch.int32).pin_memory()
            check(lib.hn_exchange_error_async(self._handle, flag.data_ptr(), stream), "hn_exchange_error_async")
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._xerr = (flag, ev)

    def check_token_sharding(self) -> None:
        """Blocking form of the check every token-sharded forward performs lazily: waits for the outstanding forward
        and raises HealNetLibraryError if a peer wait timed out."""
        if self._exchange is None:
            return
        self._poll_token_sharding(load_library(), None, wait=True)

    def _ensure_exchange(self, lib, batch: int) -> None:
        """Allocates this rank's exchange buffer, swaps CUDA IPC handles with the peers (one all-gather of 64 bytes
        per rank on the host side) and registers the peer-mapped pointers with the native handle."""
        import torch.distributed as dist
        if self._exchange is not None and self._exchange[3] >= batch and self._exchange[4] is self._handle:
            return
        if self._exchange is not None:
            # a larger batch / a new handle: unmap and free the old buffers first — once every rank has finished the
            # forwards that may still be reading them
            torch.cuda.synchronize()
            dist.barrier(group=self._exchange_group)
            self._release_exchange()
        rank, world, _ = self._token_shard
        cap = max(batch, self._exchange_max_batch)
        nbytes = lib.hn_exchange_bytes(self._handle, cap)
        mine = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        check(lib.hn_exchange_alloc(nbytes, ctypes.byref(mine), handle), "hn_exchange_alloc")
        gathered = [None] * world
        dist.all_gather_object(gathered, bytes(handle), group=self._exchange_group)
        bufs = (ctypes.c_void_p * world)()
        for r in range(world):
            if r == rank:
                bufs[r] = mine.value
            else:
                peer = ctypes.c_void_p()
                raw = (ctypes.c_ubyte * 64).from_buffer_copy(gathered[r])
                check(lib.hn_exchange_open(raw, ctypes.byref(peer)), "hn_exchange_open")
                bufs[r] = peer.value
        check(lib.hn_set_exchange(self._handle, rank, world, bufs, nbytes), "hn_set_exchange")
        check(lib.hn_set_exchange_timeout(self._handle, float(self.token_sharding_timeout_s)), "hn_set_exchange_timeout")
        dist.barrier(group=self._exchange_group)   # nobody publishes before everybody has mapped everybody
        self._exchange = (mine.value, [bufs[r] for r in range(world)], nbytes, cap, self._handle)

    def _register_attention_export(self, lib, staged, skip_self, batch, dev):
        """Allocates and registers the export buffers (hn_set_attention_export) when `export_attention_weights` is
        on; returns [(Attention module, tensor)] in call order. Shapes follow the reference: (b*h, L, N)."""
        hp = self._hparams
        M, L = self.modalities, hp["l_c"]
        if not self.export_attention_weights:
            if self._export_registered:
                for l in range(hp["depth"]):
                    for m in range(M + (1 if self.self_per_cross_attn > 0 else 0)):
                        check(lib.hn_set_attention_export(self._handle, l, m, None), "hn_set_attention_export")
                self._export_registered = False
            return []
        plan, total = [], 0
        for l, layer in enumerate(self.layers):
            last_self = None
            for m in range(M):
                if staged[m] is not None:
                    n_tok = staged[m].numel() // (batch * staged[m].shape[-1])
                    plan.append((l, m, layer[2 * m].fn, (batch * hp["x_heads"], L, n_tok)))
                if self.self_per_cross_attn > 0 and not skip_self[m]:
                    last_self = (l, M, layer[-1][0].fn, (batch * hp["l_heads"], L, L))
            if last_self is not None:
                plan.append(last_self)
        for _, _, _, shape in plan:
            total += 4 * shape[0] * shape[1] * shape[2]
        if total > self.export_attention_max_bytes:
            raise MemoryError(f"export_attention_weights needs {total / 2**30:.1f} GiB for these shapes "
                              f"(limit export_attention_max_bytes = {self.export_attention_max_bytes / 2**30:.1f} GiB)")
        out = []
        registered = set()
        for l, m, module, shape in plan:
            t = torch.empty(shape, device=dev, dtype=torch.float32)
            check(lib.hn_set_attention_export(self._handle, l, m, t.data_ptr()), "hn_set_attention_export")
            registered.add((l, m))
            out.append((module, t))
        for l in range(hp["depth"]):   # modules that do not run this time must not write into stale buffers
            for m in range(M + (1 if self.self_per_cross_attn > 0 else 0)):
                if (l, m) not in registered:
                    check(lib.hn_set_attention_export(self._handle, l, m, None), "hn_set_attention_export")
        self._export_registered = True
        return out

    # ------------------------------------------------------------------------------------------- measurement
    def capture_graph(self, tensors: Sequence[Optional[torch.Tensor]], return_embeddings: bool = False):
        """Small-batch serving: captures THIS forward (shapes, dtypes and missing modalities of `tensors`; no mask) in a
        CUDA graph and returns `run(tensors) -> output`. `run` copies the given tensors (host or device) into the
        graph's static input buffers and replays it: one launch instead of ~110, bit-identical results (the forward is
        a pure stream-ordered launch sequence; tests/test_gpu_properties.py, tools/graph_forward.py: -5 % latency at
        batch 1 on cfg 1). The returned tensor is the graph's static output buffer: consume or clone it before the next
        `run`. The graph holds the weights as they were packed at capture time: re-capture after changing them."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("capture_graph() is for inference: call it under torch.no_grad()")
        dev = next(self.parameters()).device
        static = [None if t is None else t.detach().to(dev, copy=True).contiguous() for t in tensors]
        call = lambda: self.forward(list(static), return_embeddings=return_embeddings)
        keep = self.keep_output_on_device
        self.keep_output_on_device = True
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                call()                                   # packs the weights / sizes the workspace outside the capture
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = call()
        finally:
            self.keep_output_on_device = keep
        sig = self._weights_sig

        def run(new_tensors):
            if self._weights_sig != sig:
                raise RuntimeError("the model's weights changed since capture_graph(): capture again")
            for dst, src in zip(static, new_tensors):
                if (dst is None) != (src is None) or (dst is not None and (dst.shape != src.shape or dst.dtype != src.dtype)):
                    raise ValueError("capture_graph(): inputs must keep the shapes, dtypes and missing modalities of the capture")
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
            graph.replay()
            return out

        return run

    def enable_kernel_timing(self, on: bool = True) -> None:
        """Brackets every cross-attention kernel of subsequent forwards with CUDA events (hn_profile_enable)."""
        if self._handle is None:
            dev = _compute_device(self.latents)
            with torch.cuda.device(dev):
                self._sync_native(dev, torch.cuda.current_stream(dev).cuda_stream)
        check(load_library().hn_profile_enable(self._handle, 1 if on else 0), "hn_profile_enable")

    def read_kernel_timing(self, modality: int, kind: int = 0) -> dict:
        """Device time / launches / executed and useful (unpadded) tensor FLOPs / exponentials of one class of launches
        of modality `modality` in the last forward: kind 0 = streaming cross-attention kernel, 1 = K/V projection GEMM
        (generic path), 2 = context-row build. Synchronise the stream first."""
        ms, n = ctypes.c_float(), ctypes.c_int()
        fl, fu, ex = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        check(load_library().hn_profile_read(self._handle, kind, modality, ctypes.byref(ms), ctypes.byref(n),
                                             ctypes.byref(fl), ctypes.byref(fu), ctypes.byref(ex)), "hn_profile_read")
        return dict(ms=ms.value, launches=n.value, flops=fl.value, flops_useful=fu.value, exps=ex.value)

    def get_attention_weights(self) -> List[Optional[torch.Tensor]]:
        """One entry per Attention module, in module order (healnet.py:252-262): the (b*h, L, N) softmax matrix of
        the module's last call when `export_attention_weights` was on for that forward, else None (the streaming
        kernels do not materialise it; the reference always keeps it, healnet.py:420)."""
        return [m.attn_weights for m in self.modules() if isinstance(m, Attention)]
