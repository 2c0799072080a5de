"""Drop-in ``E2VIDRecurrent`` / ``UNetRecurrent`` (config 3 of BASELINE.json) on the sm_100a kernels.

Mirrors the reference interface:

* ``E2VIDRecurrent(config: dict)``: ``forward({'events': x}) -> {'image': y}``, ``reset_states()``,
  ``num_encoders``                                          -- model/e2vid/model.py:80-116 (BaseE2VID :17-57)
* ``UNetRecurrent(...)``: ``forward(x, prev_states) -> (img, states)`` -- model/e2vid/unet.py:139-200
* ``state_dict`` keys ``unetrecurrent.{head,encoders.i.{conv,recurrent_block.Gates},resblocks.r.{conv1,conv2},
  decoders.i,pred}.*``                                    -- model/e2vid/submodules.py:16-44,78-128,212-306

One frame per call, batch is the parallel axis, recurrent state is carried between calls.  The modules are
parameter containers; all arithmetic runs in libbde2vid_sm100.so (head / encoder / decoder convs, ConvLSTM gates
conv with the pointwise update in its epilogue, ResidualBlock with the residual + ReLU in the second conv's
epilogue, bilinear x2 of (skip + x) feeding the decoder conv, fused pred + sigmoid).  No CPU path exists.
"""
import os

import torch
import torch.nn as nn

from . import ops
from .engine import VOX_CPAD, _Layer, _pack_conv
from .model import _ConvLayer, _RecurrentConv, _unsupported
from .ops import ACT_NONE, ACT_RELU, ENGINE_SIMT, ENGINE_TCGEN05, EPI_LSTM
from .registry import MODELS


class _ResidualBlock(nn.Module):                  # model/e2vid/submodules.py:212-247 (norm=None)
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1)


class LSTMState:
    """(h, c) of one encoder level.  Held in the kernels' layout (NHWC; h in the compute dtype, c in fp32);
    unpacks like the reference's ``(h, c)`` tuple of NCHW float32 tensors (model/e2vid/submodules.py:296-306)."""

    def __init__(self, h_nhwc, c_nhwc):
        self.h_nhwc, self.c_nhwc = h_nhwc, c_nhwc

    @property
    def h(self):
        return self.h_nhwc.permute(0, 3, 1, 2).float()

    @property
    def c(self):
        return self.c_nhwc.permute(0, 3, 1, 2).float()

    def __iter__(self):
        return iter((self.h, self.c))

    def __getitem__(self, i):
        return (self.h, self.c)[i]


class UNetRecurrent(nn.Module):
    """Constructor signature = reference (model/e2vid/unet.py:146-148)."""

    def __init__(self, num_bins, num_output_channels=1, skip_type='sum', recurrent_block_type='convlstm',
                 activation='sigmoid', num_encoders=4, base_num_channels=32, num_residual_blocks=2, norm=None,
                 use_upsample_conv=True):
        super().__init__()
        if skip_type != 'sum':
            _unsupported("skip_type=%r" % (skip_type,))
        if recurrent_block_type != 'convlstm':
            _unsupported("recurrent_block_type=%r" % (recurrent_block_type,))
        if activation != 'sigmoid' or num_output_channels != 1:
            _unsupported("output activation %r / %d output channels" % (activation, num_output_channels))
        if norm not in (None, 'none', 'None'):
            _unsupported("norm=%r" % (norm,))
        if not use_upsample_conv:
            _unsupported("TransposedConvLayer decoders")
        if num_bins > VOX_CPAD:
            _unsupported("num_bins > %d" % VOX_CPAD)
        self.num_bins, self.num_encoders = num_bins, num_encoders
        self.base_num_channels, self.num_residual_blocks = base_num_channels, num_residual_blocks
        bc, ne = base_num_channels, num_encoders
        self.head = _ConvLayer(num_bins, bc, 5)
        self.encoders = nn.ModuleList([_RecurrentConv(bc * 2 ** i, bc * 2 ** (i + 1), 5) for i in range(ne)])
        self.resblocks = nn.ModuleList([_ResidualBlock(bc * 2 ** ne) for _ in range(num_residual_blocks)])
        self.decoders = nn.ModuleList([_ConvLayer(bc * 2 ** (ne - i), bc * 2 ** (ne - i - 1), 5) for i in range(ne)])
        self.pred = _ConvLayer(bc, num_output_channels, 1)
        self.precision = os.environ.get("BDE2VID_PRECISION", "bf16")
        self._packed = None
        self._bufs = {}

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed, self._bufs = None, {}
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------------------------
    def _pack(self):
        from . import _lib
        _lib.require_device()
        dev = self.head.conv2d.weight.device
        if dev.type != "cuda":
            raise RuntimeError("E2VIDRecurrent (bde2vid_b200) must be on a CUDA device; no CPU path exists")
        if self.precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' (tcgen05) or 'fp32' (CUDA-core parity mode)")
        key = (self.precision, dev)
        if self._packed is not None and self._packed["key"] == key:
            return self._packed
        tc = self.precision == "bf16"
        dt = torch.bfloat16 if tc else torch.float32
        f32 = lambda t: t.detach().to(torch.float32).contiguous()  # noqa: E731

        def conv_layer(conv, stride, cin_pad=None):
            k = conv.weight.shape[-1]
            cm = tc and k > 1 and conv.weight.shape[1] % 64 == 0
            w, ld = _pack_conv(conv.weight.detach().float(), dt, cin_pad, chunk_major=cm)
            return _Layer(w, ld, f32(conv.bias), conv.weight.shape[0], k, stride, k // 2, k_order=int(cm))

        def lstm_layer(conv):   # rows reordered to n = 4*c + gate (in, remember, out, cell: e2vid/submodules.py:292)
            w = conv.weight.detach().float()
            hid = w.shape[0] // 4
            w = w.view(4, hid, *w.shape[1:]).permute(1, 0, 2, 3, 4).reshape(4 * hid, *w.shape[1:])
            b = conv.bias.detach().float().view(4, hid).t().reshape(-1)
            cm = tc and hid % 64 == 0
            pw, ld = _pack_conv(w, dt, chunk_major=cm)
            return _Layer(pw, ld, b.contiguous(), 4 * hid, 3, 1, 1, k_order=int(cm))

        with torch.no_grad():
            self._packed = dict(
                key=key, dtype=dt, engine=ENGINE_TCGEN05 if tc else ENGINE_SIMT,
                head=conv_layer(self.head.conv2d, 1, cin_pad=VOX_CPAD),
                enc=[(conv_layer(e.conv.conv2d, 2), lstm_layer(e.recurrent_block.Gates)) for e in self.encoders],
                res=[(conv_layer(r.conv1, 1), conv_layer(r.conv2, 1)) for r in self.resblocks],
                dec=[conv_layer(d.conv2d, 1) for d in self.decoders],
                pred_w=f32(self.pred.conv2d.weight.reshape(-1)), pred_b=f32(self.pred.conv2d.bias))
        return self._packed

    def _workspace(self, B, H, W, dt, dev):
        key = (B, H, W, dt)
        b = self._bufs.get(key)
        if b is not None:
            return b
        E = lambda *s, dtype=dt: torch.empty(*s, dtype=dtype, device=dev)  # noqa: E731
        bc, ne = self.base_num_channels, self.num_encoders
        b = dict(vox8=E(B, H, W, VOX_CPAD), head=E(B, H, W, bc), img=E(B, H, W, dtype=torch.float32), lv=[], dec=[])
        for i in range(ne):
            h, w, C = H >> (i + 1), W >> (i + 1), bc * 2 ** (i + 1)
            b["lv"].append(dict(h=h, w=w, C=C, e=E(B, h, w, C), zero=torch.zeros(B, h, w, C, dtype=dt, device=dev)))
        lv = b["lv"][-1]
        b["res"] = [E(B, lv["h"], lv["w"], lv["C"]) for _ in range(3)]
        for i in range(ne):
            lv = b["lv"][ne - 1 - i]
            b["dec"].append(dict(up=E(B, 2 * lv["h"], 2 * lv["w"], lv["C"]), out=E(B, 2 * lv["h"], 2 * lv["w"], lv["C"] // 2)))
        self._bufs = {key: b}
        return b

    def _gemm(self, P, layer, a0, out, n_img, h, w, c0, **kw):
        return ops.gemm(a0, layer.w, layer.bias, out, n_img=n_img, h_in=h, w_in=w, c0=c0, n=layer.n, ksize=layer.ksize,
                        stride=layer.stride, pad=layer.pad, w_ld=layer.w_ld, k_order=layer.k_order, engine=P["engine"],
                        dtype=P["dtype"], **kw)

    # ------------------------------------------------------------------------------------
    def forward(self, x, prev_states):
        """x: [N, num_bins, H, W] float32 CUDA; prev_states: None or the ``states`` of the previous call.
        Returns (img [N,1,H,W] float32, states) -- model/e2vid/unet.py:167-200."""
        if self.training:
            _unsupported("training mode")
        if not x.is_cuda:
            raise RuntimeError("E2VIDRecurrent.forward needs CUDA tensors; no CPU path exists")
        P = self._pack()
        dt = P["dtype"]
        B, bins, H, W = x.shape
        ne, bc = self.num_encoders, self.base_num_channels
        if bins != self.num_bins:
            raise ValueError("expected %d voxel bins, got %d" % (self.num_bins, bins))
        if H % (2 ** ne) or W % (2 ** ne):
            raise ValueError("input %dx%d must be padded to a multiple of %d (Croper.pad)" % (H, W, 2 ** ne))
        bufs = self._workspace(B, H, W, dt, x.device)
        if prev_states is None:
            prev_states = [None] * ne
        ops.pack_voxel_nhwc(x.to(torch.float32).contiguous(), VOX_CPAD, dt, out=bufs["vox8"])
        self._gemm(P, P["head"], bufs["vox8"], bufs["head"], B, H, W, VOX_CPAD, act=ACT_RELU)
        cur, cc, ch, cw = bufs["head"], bc, H, W
        states = []
        for i in range(ne):
            lv = bufs["lv"][i]
            conv, lstm = P["enc"][i]
            self._gemm(P, conv, cur, lv["e"], B, ch, cw, cc, act=ACT_RELU)
            st = prev_states[i]
            if st is not None and not isinstance(st, LSTMState):      # reference-layout (h, c) NCHW tensors
                h0, c0 = st
                st = LSTMState(h0.permute(0, 2, 3, 1).to(dt).contiguous(), c0.permute(0, 2, 3, 1).float().contiguous())
            # states are values (the caller may keep them): fresh tensors per step, as in the reference
            h_out = torch.empty(B, lv["h"], lv["w"], lv["C"], dtype=dt, device=x.device)
            c_out = torch.empty(B, lv["h"], lv["w"], lv["C"], dtype=torch.float32, device=x.device)
            self._gemm(P, lstm, lv["e"], h_out, B, lv["h"], lv["w"], lv["C"],
                       a1=lv["zero"] if st is None else st.h_nhwc, c1=lv["C"], epi=EPI_LSTM,
                       c_prev=None if st is None else st.c_nhwc, c_out=c_out)
            states.append(LSTMState(h_out, c_out))
            cur, cc, ch, cw = h_out, lv["C"], lv["h"], lv["w"]
        # residual blocks: relu(conv2(relu(conv1(x))) + x)   (e2vid/submodules.py:234-247)
        r = bufs["res"]
        for j, (c1l, c2l) in enumerate(P["res"]):
            self._gemm(P, c1l, cur, r[2], B, ch, cw, cc, act=ACT_RELU)
            dst = r[j & 1]
            self._gemm(P, c2l, r[2], dst, B, ch, cw, cc, act=ACT_RELU, residual=cur, res_mode=1)
            cur = dst
        # decoders: UpsampleConvLayer(skip_sum(x, block))   (unet.py:193-195; ReLU, e2vid/submodules.py:78-106)
        for i in range(ne):
            lv = bufs["lv"][ne - 1 - i]
            dd = bufs["dec"][i]
            blk = states[ne - 1 - i].h_nhwc
            ops.upsample2x_sum(blk, cur, 1.0, B, lv["h"], lv["w"], lv["C"], dd["up"])
            self._gemm(P, P["dec"][i], dd["up"], dd["out"], B, 2 * lv["h"], 2 * lv["w"], lv["C"], act=ACT_RELU)
            cur = dd["out"]
        ops.pred_sigmoid(cur, bufs["head"], P["pred_w"], P["pred_b"], bc, B * H * W, bufs["img"])
        return bufs["img"].clone().view(B, 1, H, W), states


@MODELS.register_module()
class E2VIDRecurrent(nn.Module):
    """Drop-in for model/e2vid/model.py:80-116 (config parsing: BaseE2VID :17-57)."""

    def __init__(self, config):
        super().__init__()
        assert 'num_bins' in config
        self.num_bins = int(config['num_bins'])
        self.skip_type = str(config.get('skip_type', 'sum'))
        self.num_encoders = int(config.get('num_encoders', 4))
        self.base_num_channels = int(config.get('base_num_channels', 32))
        self.num_residual_blocks = int(config.get('num_residual_blocks', 2))
        self.norm = config.get('norm', None)
        self.use_upsample_conv = bool(config.get('use_upsample_conv', True))
        self.kernel_size = int(config.get('kernel_size', 5))
        self.recurrent_block_type = str(config.get('recurrent_block_type', 'convlstm'))
        self.unetrecurrent = UNetRecurrent(num_bins=self.num_bins, num_output_channels=1, skip_type=self.skip_type,
                                           recurrent_block_type=self.recurrent_block_type, activation='sigmoid',
                                           num_encoders=self.num_encoders, base_num_channels=self.base_num_channels,
                                           num_residual_blocks=self.num_residual_blocks, norm=self.norm,
                                           use_upsample_conv=self.use_upsample_conv)
        self.prev_states = None

    def reset_states(self):
        self.prev_states = None

    def forward(self, inputs):
        img_pred, self.prev_states = self.unetrecurrent.forward(inputs['events'], self.prev_states)
        return {'image': img_pred}
