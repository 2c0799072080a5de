"""Drop-in ``BDE2VID`` model class backed by libbde2vid_sm100.so.

Mirrors the reference interface for the inference hot path:

* ``BDE2VID(generator: dict, cpu_cache_length=100, init_cfg=None)``, ``forward(inputs, mode='tensor')``,
  ``reset_states()``                                    -- model/BDE2VID/bde2vid.py:12-50
* ``BDE2VIDCrossscalePropogationV5(**cfg)`` generator with the reference's constructor arguments and
  ``state_dict`` key layout                            -- model/BDE2VID/bde2vid_cross_scale_propogation_V5.py:17-98
* both registered in ``MODELS`` so that ``MODELS.build(cfg.model)`` + ``load_state_dict`` of a
  reference checkpoint works (eval_models_seq.py:52-60, :86).

The nn.Modules below are parameter containers only (they give ``load_state_dict``/``state_dict``
the reference's key names).  ``forward`` never touches torch ops for arithmetic: it drives the
hand-written kernels through the C ABI; PyTorch supplies device memory and streams.
Schedule and quirks Q1-Q4 follow ...V5.py:100-241.
"""
import os

import torch
import torch.nn as nn

from . import _lib, ops
from .registry import MODELS

# --------------------------------------------------------------------------------------------
# parameter containers with the reference's attribute names
# --------------------------------------------------------------------------------------------


def _norm_layer(norm, c):
    """ConvLayer / UpsampleConvLayer norm (submodules.py:100-103, 133-136): eval-mode BatchNorm / InstanceNorm with
    running statistics, folded into the convolution weights when the engine packs them."""
    if norm == 'BN':
        return nn.BatchNorm2d(c)
    if norm == 'IN':
        return nn.InstanceNorm2d(c, track_running_stats=True)
    return None


class _ConvLayer(nn.Module):                      # submodules.py:85  (conv2d [+ norm_layer])
    def __init__(self, cin, cout, k, norm=None):
        super().__init__()
        self.conv2d = nn.Conv2d(cin, cout, k, padding=k // 2, bias=norm != 'BN')
        nl = _norm_layer(norm, cout)
        if nl is not None:
            self.norm_layer = nl


class _ConvLSTM(nn.Module):                       # submodules.py:278 (Gates)
    def __init__(self, cin, hidden, k=3):
        super().__init__()
        self.Gates = nn.Conv2d(cin + hidden, 4 * hidden, k, padding=k // 2)


class _ConvGRU(nn.Module):                        # submodules.py:337 (reset_gate, update_gate, out_gate)
    def __init__(self, cin, hidden, k=3):
        super().__init__()
        self.reset_gate = nn.Conv2d(cin + hidden, hidden, k, padding=k // 2)
        self.update_gate = nn.Conv2d(cin + hidden, hidden, k, padding=k // 2)
        self.out_gate = nn.Conv2d(cin + hidden, hidden, k, padding=k // 2)


class _RecurrentConv(nn.Module):                  # submodules.py:173 (conv, recurrent_block)
    def __init__(self, cin, cout, k, norm=None, recurrent_block_type='convlstm'):
        super().__init__()
        self.conv = _ConvLayer(cin, cout, k, norm)
        self.recurrent_block = (_ConvLSTM if recurrent_block_type == 'convlstm' else _ConvGRU)(cout, cout, 3)


class _WindowAttention(nn.Module):                # DTransformer.py:99
    def __init__(self, dim, D, wh, ww, heads, nwin=None):
        super().__init__()
        from .synth import relative_position_index
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * D - 1) * (2 * wh - 1) * (2 * ww - 1), heads))
        self.register_buffer("relative_position_index", relative_position_index(D, wh, ww))
        self.norm_q = nn.LayerNorm(dim)
        self.norm_kv = nn.LayerNorm(dim)
        if nwin is not None:                      # "feature reduction" (DTransformer.py:128-131)
            self.reduction_conv = nn.Conv2d(dim, nwin[0] * nwin[1] * dim, kernel_size=(wh, ww), groups=dim)
        self.q = nn.Linear(dim, dim)
        self.kv = nn.Linear(dim, 2 * dim)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class _Mlp(nn.Module):                            # DTransformer.py:19
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.fc2 = nn.Linear(4 * dim, dim)


class _SwinBlock(nn.Module):                      # DTransformer.py:213
    def __init__(self, dim, D, wh, ww, heads, nwin=None):
        super().__init__()
        self.attn = _WindowAttention(dim, D, wh, ww, heads, nwin)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim)


class _DFrameAttention(nn.Module):                # DTransformer.py:309
    def __init__(self, dim, depth, D, wh, ww, heads, nwin=None):
        super().__init__()
        self.blocks = nn.ModuleList([_SwinBlock(dim, D, wh, ww, heads, nwin) for _ in range(depth)])


class _ResidualBlockNoBN(nn.Module):              # ...V5.py:261-275 (conv1, conv2)
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1)


def _unsupported(what):
    raise NotImplementedError("bde2vid_b200: %s is not implemented in the sm_100a path "
                              "(there is no PyTorch/CPU fallback)" % what)


_OUT_ACTS = {"Sigmoid": ops.ACT_SIGMOID, "Identity": ops.ACT_NONE}
_NET_ACTS = {"default": ops.ACT_RELU, "ReLU": ops.ACT_RELU, "ReLU6": ops.ACT_RELU6}


@MODELS.register_module()
class BDE2VIDCrossscalePropogationV5(nn.Module):
    """Generator.  Constructor signature = reference (...V5.py:19-23).  Every architecture option of the reference
    constructor is built from the same kernels: ``norm`` None / 'BN' / 'IN' (eval-mode statistics folded into the
    convolution), ``recurrent_block_type`` 'convlstm' / 'convgru', ``useRC``, ``skip_type`` 'sum' / 'concat',
    ``nwindow_size`` (reduction_conv), ``depths[-1] == 0`` (ParseLayer + ResidualBlockNoBN tail), output activation
    'Sigmoid' / 'Identity'."""

    def __init__(self, num_bins, basechannels, num_encoders, ks, num_res_blocks, norm=None,
                 recurrent_block_type='convlstm', useRC=True, skip_type='sum', activation=None,
                 num_output_channels=1, act_net="default", buffer_index=None, q_idx=None,
                 window_size=(7, 7), nwindow_size=None, depths=[4, 0, 6], num_heads=16, drop_path_rate=0.2,
                 use_checkpoint=False, act_attn="default", losses=None, loss_inds=None, init_cfg=None):
        super().__init__()
        if norm in ("none", "None"):
            norm = None
        if norm not in (None, "BN", "IN"):
            _unsupported("norm=%r" % (norm,))
        if recurrent_block_type not in ('convlstm', 'convgru'):
            raise AssertionError("recurrent_block_type must be 'convlstm' or 'convgru' (submodules.py:179)")
        if skip_type in ('no_skip', None):
            _unsupported("skip_type=%r (the reference itself fails on it: nn.Identity() returns the [skip, x] list, "
                         "...V5.py:32-33,194)" % (skip_type,))
        if skip_type not in ('sum', 'concat'):
            raise KeyError('Could not identify skip_type, please add "skip_type": "sum", "concat" or "no_skip" '
                           'to config["model"]')
        act_name = "Sigmoid" if activation is None else dict(activation).get("type", "Sigmoid")
        if act_name not in _OUT_ACTS:
            _unsupported("output activation %r" % (activation,))
        if act_net not in _NET_ACTS or act_attn not in ("default", "GELU"):
            _unsupported("act_net=%r / act_attn=%r" % (act_net, act_attn))
        if num_output_channels != 1:
            _unsupported("num_output_channels != 1")
        if buffer_index is None or q_idx is None:
            raise ValueError("buffer_index and q_idx are required (they have no default in the reference either)")
        depths = list(depths)
        if len(depths) != num_encoders:
            raise ValueError("len(depths) must equal num_encoders")
        if len(buffer_index) > 8:
            _unsupported("more than 8 buffered frames")
        nwin = None if nwindow_size is None else tuple(int(v) for v in nwindow_size)
        self.cfg = dict(num_bins=num_bins, basechannels=basechannels, num_encoders=num_encoders, ks=ks,
                        buffer_index=[int(b) for b in buffer_index], q_idx=int(q_idx),
                        window_size=tuple(window_size), depths=depths, num_heads=num_heads, norm=norm,
                        recurrent_block_type=recurrent_block_type if useRC else None, skip_type=skip_type,
                        nwindow_size=nwin, num_res_blocks=num_res_blocks, out_act=_OUT_ACTS[act_name],
                        net_act=_NET_ACTS[act_net])
        self.losses_cfg = losses          # accepted and ignored: training-only (...V5.py:37-38)
        self.num_encoders = num_encoders
        bc, ne = basechannels, num_encoders
        D = len(buffer_index)
        wh, ww = tuple(window_size)
        self.head = _ConvLayer(num_bins, bc, ks, norm)

        def encoder():                            # Encoder(), ...V5.py:244-259
            if useRC:
                return nn.ModuleList([_RecurrentConv(bc * 2 ** i, bc * 2 ** (i + 1), ks, norm, recurrent_block_type)
                                      for i in range(ne)])
            return nn.ModuleList([_ConvLayer(bc * 2 ** i, bc * 2 ** (i + 1), ks, norm) for i in range(ne)])

        self.forward_encoder = encoder()
        self.backward_encoder = encoder()
        # present in every reference checkpoint, never used by forward (quirk Q3)
        self.fusion_layers = nn.ModuleList([nn.Conv2d(bc * 2 ** (i + 2), bc * 2 ** (i + 1), 1) for i in range(ne)])
        self.feat_attns = nn.ModuleList([
            _DFrameAttention(bc * 2 ** (i + 1), d, D, wh, ww, num_heads, nwin) if d > 0 else None
            for i, d in enumerate(depths)])
        if depths[-1] == 0:                       # ...V5.py:77-80: ParseLayer + ResidualBlockNoBN x num_res_blocks
            self.feat_attns[-1] = nn.Sequential(nn.Identity(), *[_ResidualBlockNoBN(bc * 2 ** ne) for _ in range(num_res_blocks)])
        concat = skip_type == 'concat'
        self.decoders = nn.ModuleList([
            nn.Sequential(nn.Conv2d(bc * 2 ** (ne - i + 1), bc * 2 ** (ne - i), 1) if concat else nn.Identity(),
                          _ConvLayer(bc * 2 ** (ne - i), bc * 2 ** (ne - i - 1), ks, norm)) for i in range(ne)])
        self.predI = nn.Sequential(nn.Conv2d(bc * 2, bc, 1) if concat else nn.Identity(),
                                   nn.Conv2d(bc, num_output_channels, 1))
        self._engine = None
        self.precision = os.environ.get("BDE2VID_PRECISION", "bf16")
        self.use_cuda_graph = os.environ.get("BDE2VID_CUDA_GRAPH", "1") != "0"

    # any weight change invalidates the packed copies
    def _load_from_state_dict(self, *a, **k):
        self._engine = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self):
        from .engine import Engine
        dev = self.head.conv2d.weight.device
        if self._engine is None or self._engine.precision != self.precision or self._engine.device != dev:
            self._engine = Engine(self, self.precision)
        return self._engine

    def forward(self, input_seqs, record=False, out_preds=True, out_loss=False, cpu_cache_length=100):
        if out_loss or record:
            _unsupported("loss / record modes")
        if self.training:
            _unsupported("training mode")
        vox = [d['events'] for d in input_seqs]
        predicts = self.engine().forward(vox, use_graph=self.use_cuda_graph)
        return None, (predicts if out_preds else None), None, None, None

    def reset_states(self):
        pass  # recurrent state lives in per-call buffers; forward() always starts from zeros (bde2vid.py:31)


@MODELS.register_module()
class BDE2VID(nn.Module):
    """Drop-in for model/BDE2VID/bde2vid.py:12-50."""

    def __init__(self, generator, cpu_cache_length=100, init_cfg=None):
        super().__init__()
        self.cpu_cache_length = cpu_cache_length   # kept for API compatibility; nothing is offloaded
        self.generator_cfg = generator
        self.generator = MODELS.build(generator)
        self.vis = None

    def reset_states(self):
        self.generator.reset_states()
        if not self.training:
            self.vis = None

    def forward(self, inputs, mode='tensor', **kwargs):
        self.reset_states()
        if mode == 'tensor':
            _, predicts, _, _, _ = self.generator(inputs, record=False, out_preds=True, out_loss=False,
                                                  cpu_cache_length=self.cpu_cache_length)
            return predicts
        _unsupported("forward mode %r (only 'tensor' is on the inference path)" % (mode,))

    # fused entry point: raw events -> frames without materialising voxel grids on the host
    def reconstruct_events(self, xs, ys, ts, ps, offsets, sensor_size, num_encoders=None, slot=0, normalize=None,
                           hot_mask=None):
        """events + CSR offsets (CUDA or pinned host) -> list of T cropped frames [1,1,H,W].  Events are either the
        loader format (four float32 arrays, h5_dataset.py:222-225) or the on-disk dtypes (int16 x / y, float64 t, bool p:
        event_packagers.py:44-47), converted in the voxeliser.  The loader contract is kept: windows with fewer than 3
        events give zero grids (h5_dataset.py:219-221); ``normalize`` ('legacy' | 'robust' | ('robust', low, top)) and
        ``hot_mask`` apply the loader's voxel transforms (h5_dataset.py:226, :364) on the device.  ``slot`` picks an
        independent buffer set so that several sequences can run concurrently on different CUDA streams."""
        from .croper import Croper
        H, W = sensor_size
        crop = Croper(self.generator.num_encoders if num_encoders is None else num_encoders)
        crop.update_params(W, H)
        eng = self.generator.engine()
        frames = eng.forward_events(xs, ys, ts, ps, offsets, H, W, crop, use_graph=self.generator.use_cuda_graph,
                                    slot=slot, normalize=normalize, hot_mask=hot_mask)
        return [crop.crop(f) for f in frames]

    def reconstruct_events_batch(self, seqs, sensor_size, num_encoders=None, slot=0, normalize=None, hot_mask=None):
        """Several independent event sequences (same window count) reconstructed as ONE batch: ``seqs`` is a
        list of (xs, ys, ts, ps, offsets) tuples; returns a list (per sequence) of T cropped frames."""
        from .croper import Croper
        H, W = sensor_size
        crop = Croper(self.generator.num_encoders if num_encoders is None else num_encoders)
        crop.update_params(W, H)
        eng = self.generator.engine()
        out = eng.forward_events_batch(seqs, H, W, crop, use_graph=self.generator.use_cuda_graph, slot=slot,
                                       normalize=normalize, hot_mask=hot_mask)
        return [[crop.crop(f) for f in frames] for frames in out]


    def reconstruct_events_pair(self, xs, ys, ts, ps, offsets, sensor_size, pair_rank, group=None, src_ranks=(0, 1),
                                num_encoders=None, normalize=None, hot_mask=None, graph=False):
        """ONE sequence split over TWO GPUs: both ranks of the pair (process group ``group``,
        global ranks ``src_ranks``) call this with the same events; rank 0 runs the forward recurrent chains, rank 1 the
        backward ones, the hidden states are exchanged per level over NCCL, decoder chunks alternate, and both ranks return
        all T frames (bit-identical to ``reconstruct_events`` on one GPU).  See ``engine.PairSplit``."""
        from .croper import Croper
        from .engine import PairSplit
        H, W = sensor_size
        crop = Croper(self.generator.num_encoders if num_encoders is None else num_encoders)
        crop.update_params(W, H)
        eng = self.generator.engine()
        split = PairSplit(pair_rank, group, src_ranks, graph)
        out = eng.forward_events_batch([(xs, ys, ts, ps, offsets)], H, W, crop, use_graph=self.generator.use_cuda_graph,
                                       normalize=normalize, hot_mask=hot_mask, pair=split)
        return [crop.crop(f) for f in out[0]]


def load_checkpoint(path_or_dict, device="cuda"):
    """Reference loading path (eval_models_seq.py:41-60,:86) for mmengine-style checkpoints."""
    from .registry import Config
    ckpt = torch.load(path_or_dict, map_location="cpu", weights_only=False) if isinstance(path_or_dict, str) else path_or_dict
    cfg = Config.fromstring(ckpt['meta']['cfg'], '.py').model
    model = MODELS.build(cfg)
    model.load_state_dict(ckpt['state_dict'])
    return model.eval().to(device)
