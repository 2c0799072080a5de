"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this file; the product package ``bde2vid_b200`` never does.

What is restated (all citations are into the reference tree, gaopinghai/BDE2VID):

* ``voxel_grid``                -- events_contrast_maximization/utils/event_utils.py:466-509
                                   (+ events_to_image_torch :330-376)
* ``croper_params``             -- utils_func/inference_utils.py:26-32, 83-114
* ``bde2vid_forward``           -- model/BDE2VID/bde2vid_cross_scale_propogation_V5.py:100-241
  (level-by-level bidirectional schedule incl. quirks Q1..Q4 of SURVEY.md section 3.2)
* ``convlstm_step``             -- model/BDE2VID/submodules.py:278-334
* ``upsample_conv``             -- model/BDE2VID/submodules.py:117-147
* ``attention_block`` / ``dframe_attention`` -- model/BDE2VID/DTransformer.py:99-389
* ``e2vid_recurrent_forward``   -- model/e2vid/unet.py:139-200, model/e2vid/submodules.py

The functions work on a plain ``state_dict`` with the reference's key names, in fp32 on the
CPU, with functional torch ops (conv2d / linear / softmax ... live in PyTorch, a third-party
dependency of the reference with no pinned version; here torch 2.11).

Pinning: the reference has no tests or golden vectors.  This port is
pinned against outputs of the reference itself, run in the build container through
``oracle/ref_shim.py`` -- see ``tests/test_oracle_vs_reference.py`` (live, container only)
and the committed fixtures in ``tests/golden`` written by ``oracle/make_golden.py``.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------

DEFAULT_CFG = dict(
    type="BDE2VIDCrossscalePropogationV5",
    num_bins=5, basechannels=32, num_encoders=3, ks=5, num_res_blocks=2, norm=None,
    recurrent_block_type="convlstm", useRC=True, skip_type="sum",
    activation=dict(type="Sigmoid"), num_output_channels=1, act_net="default",
    buffer_index=[-1, 0, 1], q_idx=1, window_size=(7, 7), nwindow_size=None,
    depths=[4, 0, 6], num_heads=16, losses=[],
)
# keys of DEFAULT_CFG that are constructor arguments of the generator (everything but 'type')
GEN_KEYS = [k for k in DEFAULT_CFG if k != "type"]


def full_cfg(cfg=None):
    out = dict(DEFAULT_CFG)
    if cfg:
        out.update(cfg)
    return out


# --------------------------------------------------------------------------------------
# voxeliser  (event_utils.py:466-509, :330-376)
# --------------------------------------------------------------------------------------

def voxel_grid(xs, ys, ts, ps, num_bins, sensor_size):
    """Temporal-bilinear voxel grid, numpy fp32, events accumulated in index order.

    event_utils.py:489-490  t_norm = (ts - ts[0]) / dt * (B-1)         (fp32, that op order)
    event_utils.py:494-495  w_b = ps * max(0, 1 - |t_norm - b|)        (for every bin b)
    event_utils.py:371-375  img[int(y), int(x)] += w_b                 (truncating cast)
    """
    xs = np.asarray(xs, dtype=np.float32)
    ys = np.asarray(ys, dtype=np.float32)
    ts = np.asarray(ts, dtype=np.float32)
    ps = np.asarray(ps, dtype=np.float32)
    assert len(xs) == len(ys) == len(ts) == len(ps)
    H, W = sensor_size
    dt = np.float32(ts[-1] - ts[0])
    t_norm = ((ts - ts[0]) / dt * np.float32(num_bins - 1)).astype(np.float32)
    xi = xs.astype(np.int64)
    yi = ys.astype(np.int64)
    grid = np.zeros((num_bins, H, W), dtype=np.float32)
    one = np.float32(1.0)
    for b in range(num_bins):
        w = (ps * np.maximum(np.float32(0.0), one - np.abs(t_norm - np.float32(b)))).astype(np.float32)
        np.add.at(grid[b], (yi, xi), w)
    return grid


def voxel_bin_indices(ts, num_bins):
    """Left bin index floor(t_norm) of every event (the bit-exact gate of the north star)."""
    ts = np.asarray(ts, dtype=np.float32)
    dt = np.float32(ts[-1] - ts[0])
    t_norm = ((ts - ts[0]) / dt * np.float32(num_bins - 1)).astype(np.float32)
    return np.floor(t_norm).astype(np.int32), t_norm


def voxel_abs_mass(xs, ys, ts, ps, num_bins, sensor_size):
    """Sum of |contribution| per cell: the scale of the well-posed value gate."""
    return voxel_grid(xs, ys, ts, np.abs(np.asarray(ps, dtype=np.float32)), num_bins, sensor_size)


# --------------------------------------------------------------------------------------
# pad / crop  (inference_utils.py:26-32, 83-114)
# --------------------------------------------------------------------------------------

def croper_params(width, height, num_encoders):
    S = 2 ** num_encoders
    Wp = S * math.ceil(width / S)
    Hp = S * math.ceil(height / S)
    pt, pb = math.ceil(0.5 * (Hp - height)), math.floor(0.5 * (Hp - height))
    pl, pr = math.ceil(0.5 * (Wp - width)), math.floor(0.5 * (Wp - width))
    cx, cy = Wp // 2, Hp // 2
    ix0, ix1 = cx - width // 2, cx + math.ceil(width / 2)
    iy0, iy1 = cy - height // 2, cy + math.ceil(height / 2)
    return dict(Hp=Hp, Wp=Wp, pad=(pl, pr, pt, pb), crop=(iy0, iy1, ix0, ix1))


def pad_voxel(v, params):
    return F.pad(v, params["pad"])


def crop_image(img, params):
    iy0, iy1, ix0, ix1 = params["crop"]
    return img[..., iy0:iy1, ix0:ix1]


# --------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------

def _act(x, name):
    if name in (None, "none"):
        return x
    if name in ("ReLU", "relu", "default"):
        return F.relu(x)
    if name == "ReLU6":
        return F.relu6(x)
    raise NotImplementedError(name)


def _norm(sd, prefix, x):
    """norm_layer of ConvLayer / UpsampleConvLayer in eval mode (submodules.py:100-111, 133-145): BatchNorm2d, or
    InstanceNorm2d(track_running_stats=True) which then also normalises with its running statistics (no affine)."""
    if prefix + ".norm_layer.running_mean" not in sd:
        return x
    return F.batch_norm(x, sd[prefix + ".norm_layer.running_mean"], sd[prefix + ".norm_layer.running_var"],
                        sd.get(prefix + ".norm_layer.weight"), sd.get(prefix + ".norm_layer.bias"), False, 0.1, 1e-5)


def conv_layer(sd, prefix, x, stride, act):
    """ConvLayer (submodules.py:85-114): conv2d (no bias with BN) -> optional norm -> activation."""
    w = sd[prefix + ".conv2d.weight"]
    y = F.conv2d(x, w, sd.get(prefix + ".conv2d.bias"), stride=stride, padding=w.shape[-1] // 2)
    return _act(_norm(sd, prefix, y), act)


def convlstm_step(sd, prefix, x, state):
    """ConvLSTM (submodules.py:316-332): chunk order in, remember, out, cell."""
    w = sd[prefix + ".Gates.weight"]
    hid = w.shape[0] // 4
    if state is None:
        z = torch.zeros(x.shape[0], hid, x.shape[2], x.shape[3], dtype=x.dtype)
        state = (z, z)
    h_prev, c_prev = state
    g = F.conv2d(torch.cat([x, h_prev], 1), w, sd[prefix + ".Gates.bias"], padding=w.shape[-1] // 2)
    gi, gf, go, gc = g.chunk(4, 1)
    c = torch.sigmoid(gf) * c_prev + torch.sigmoid(gi) * torch.tanh(gc)
    h = torch.sigmoid(go) * torch.tanh(c)
    return h, c


def convgru_step(sd, prefix, x, state):
    """ConvGRU (submodules.py:355-375; model/submodules.py and model/e2vid/submodules.py hold the same cell)."""
    wu = sd[prefix + ".update_gate.weight"]
    pad = wu.shape[-1] // 2
    if state is None:
        state = torch.zeros(x.shape[0], wu.shape[0], x.shape[2], x.shape[3], dtype=x.dtype)
    stacked = torch.cat([x, state], 1)
    update = torch.sigmoid(F.conv2d(stacked, wu, sd[prefix + ".update_gate.bias"], padding=pad))
    reset = torch.sigmoid(F.conv2d(stacked, sd[prefix + ".reset_gate.weight"], sd[prefix + ".reset_gate.bias"], padding=pad))
    out = torch.tanh(F.conv2d(torch.cat([x, state * reset], 1), sd[prefix + ".out_gate.weight"],
                              sd[prefix + ".out_gate.bias"], padding=pad))
    return state * (1 - update) + out * update


def recurrent_step(sd, prefix, x, state):
    """recurrent_block of a RecurrentConv / RecurrentConvLayer: returns (output, new state)."""
    if prefix + ".Gates.weight" in sd:
        st = convlstm_step(sd, prefix, x, state)
        return st[0], st
    st = convgru_step(sd, prefix, x, state)
    return st, st


def upsample_conv(sd, prefix, x, act):
    """UpsampleConvLayer (submodules.py:137-147): bilinear x2 (align_corners=False), conv, optional norm, activation."""
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    w = sd[prefix + ".conv2d.weight"]
    y = F.conv2d(x, w, sd.get(prefix + ".conv2d.bias"), padding=w.shape[-1] // 2)
    return _act(_norm(sd, prefix, y), act)


# --------------------------------------------------------------------------------------
# window attention  (DTransformer.py)
# --------------------------------------------------------------------------------------

def window_geometry(H, W, ws=(7, 7)):
    """Pad amounts and window grid of one attention block (DTransformer.py:254-266)."""
    wh = H if H <= ws[0] else ws[0]
    ww = W if W <= ws[1] else ws[1]
    pad_h = (wh - H % wh) % wh
    pad_w = (ww - W % ww) % ww
    pt, pl = pad_h // 2, pad_w // 2
    Hg, Wg = H + pad_h, W + pad_w
    return dict(wh=wh, ww=ww, pt=pt, pl=pl, Hg=Hg, Wg=Wg, nH=Hg // wh, nW=Wg // ww)


def window_token_map(H, W, dilated, ws=(7, 7)):
    """For every window n and token (a, b): source pixel in the *unpadded* map, or -1 if the
    token falls on padding.  Restates window_partition (DTransformer.py:41-60):
    plain  -> pixel (wh*i + a, ww*j + b) of the symmetric-padded map;
    dilated-> pixel (wh*i + 2a, ww*j + 2b) of that map padded again by (wh, ww) bottom/right
              (F.unfold kernel=window, dilation 2, stride=window)."""
    g = window_geometry(H, W, ws)
    wh, ww, nH, nW = g["wh"], g["ww"], g["nH"], g["nW"]
    step = 2 if dilated else 1
    i = torch.arange(nH).view(nH, 1, 1, 1)
    j = torch.arange(nW).view(1, nW, 1, 1)
    a = torch.arange(wh).view(1, 1, wh, 1)
    b = torch.arange(ww).view(1, 1, 1, ww)
    r = wh * i + step * a - g["pt"]        # row in the unpadded map
    c = ww * j + step * b - g["pl"]
    r = r.expand(nH, nW, wh, ww)
    c = c.expand(nH, nW, wh, ww)
    ok = (r >= 0) & (r < H) & (c >= 0) & (c < W)
    # rows/cols inside the padded grid but outside the real map are zero tokens; rows beyond
    # the grid (dilated extra pad) are zero tokens too -- both map to -1.
    idx = torch.where(ok, r * W + c, torch.full_like(r, -1))
    # destination (for the reverse scatter): pixel in the padded grid, must be < Hg/Wg to survive the crop
    rg = wh * i + step * a
    cg = ww * j + step * b
    rg = rg.expand(nH, nW, wh, ww)
    cg = cg.expand(nH, nW, wh, ww)
    return idx.reshape(nH * nW, wh * ww), ok.reshape(nH * nW, wh * ww), g


def rel_pos_bias(sd, prefix, D, wh, ww, q_ind, num_heads, n_kv=None):
    """relative_position_bias_table gathered for the query frame's tokens
    (DTransformer.py:139-152, 195-199) -> [nH, wh*ww, N]; N = D*wh*ww, or the first N columns with nwindow_size."""
    table = sd[prefix + ".relative_position_bias_table"]
    index = sd[prefix + ".relative_position_index"]
    n = wh * ww
    N = D * n if n_kv is None else n_kv
    idx = index[q_ind * n:(q_ind + 1) * n, :N].reshape(-1)
    return table[idx].reshape(n, N, num_heads).permute(2, 0, 1).contiguous()


def attention_block(sd, prefix, frames, q_ind, num_heads, dilated, ws=(7, 7)):
    """SwinTransformerBlock3D.forward (DTransformer.py:285-306) on ``frames`` = list of D maps
    [B, C, H, W]; returns the updated query-frame map."""
    D = len(frames)
    B, C, H, W = frames[0].shape
    idx, ok, g = window_token_map(H, W, dilated, ws)
    nWin, n = idx.shape
    hd = C // num_heads
    ap = prefix + ".attn"

    def tokens(fr):  # [B, C, H, W] -> [B, nWin, n, C] with zero tokens on padding
        flat = fr.permute(0, 2, 3, 1).reshape(B, H * W, C)
        t = flat[:, idx.clamp(min=0).reshape(-1)].reshape(B, nWin, n, C)
        return t * ok.view(1, nWin, n, 1).to(t.dtype)

    tok = [tokens(fr) for fr in frames]
    q_in = tok[q_ind]
    if ap + ".reduction_conv.weight" in sd:
        # "feature reduction" (DTransformer.py:128-131, 172-175): depthwise whole-window conv Conv2d(C, X*C, window, groups=C)
        # on every (frame, window); the [C*X, 1, 1] result is VIEWED as [X, C] (index j*C + c), as the reference does
        rw, rb = sd[ap + ".reduction_conv.weight"], sd[ap + ".reduction_conv.bias"]
        X = rw.shape[0] // C
        # one conv over all D frames' windows, contiguous NCHW like the reference's x.view(-1, C, H, W) (same batch layout,
        # so that the CPU conv takes the same code path and the port stays bit-equal)
        img = torch.stack([t.reshape(B * nWin, g["wh"], g["ww"], C).permute(0, 3, 1, 2) for t in tok], 0).contiguous()
        red = F.conv2d(img.view(-1, C, g["wh"], g["ww"]), rw, rb, groups=C).view(D, B, nWin, X, C)
        kv_in = red.permute(1, 2, 0, 3, 4).reshape(B, nWin, D * X, C)   # [B, nWin, D*X, C], index d*X + j
        n_kv = D * X
    else:
        kv_in = torch.cat(tok, dim=2)                              # [B, nWin, D*n, C], index d*n + a*ww + b
        n_kv = D * n
    qn = F.layer_norm(q_in, (C,), sd[ap + ".norm_q.weight"], sd[ap + ".norm_q.bias"], 1e-5)
    kvn = F.layer_norm(kv_in, (C,), sd[ap + ".norm_kv.weight"], sd[ap + ".norm_kv.bias"], 1e-5)
    q = F.linear(qn, sd[ap + ".q.weight"], sd[ap + ".q.bias"])
    kv = F.linear(kvn, sd[ap + ".kv.weight"], sd[ap + ".kv.bias"])
    q = q.reshape(B, nWin, n, num_heads, hd).permute(0, 1, 3, 2, 4) * (hd ** -0.5)
    k = kv[..., :C].reshape(B, nWin, n_kv, num_heads, hd).permute(0, 1, 3, 2, 4)
    v = kv[..., C:].reshape(B, nWin, n_kv, num_heads, hd).permute(0, 1, 3, 2, 4)
    attn = q @ k.transpose(-2, -1)
    attn = attn + rel_pos_bias(sd, ap, D, g["wh"], g["ww"], q_ind, num_heads, n_kv).view(1, 1, num_heads, n, n_kv)
    attn = torch.softmax(attn, dim=-1)
    o = (attn @ v).permute(0, 1, 3, 2, 4).reshape(B, nWin, n, C)
    o = F.linear(o, sd[ap + ".proj.weight"], sd[ap + ".proj.bias"])
    # window_reverse + crop: every real pixel is covered by at most one window token; pixels
    # covered by none (dilated blocks: grid rows/cols 1,3,5) receive 0 (F.fold semantics).
    out = torch.zeros(B, H * W, C, dtype=o.dtype)
    sel = ok.reshape(-1)
    out[:, idx.reshape(-1)[sel]] = o.reshape(B, nWin * n, C)[:, sel]
    x = frames[q_ind].permute(0, 2, 3, 1).reshape(B, H * W, C) + out
    y = F.layer_norm(x, (C,), sd[prefix + ".norm2.weight"], sd[prefix + ".norm2.bias"], 1e-5)
    y = F.linear(y, sd[prefix + ".mlp.fc1.weight"], sd[prefix + ".mlp.fc1.bias"])
    y = F.gelu(y)
    y = F.linear(y, sd[prefix + ".mlp.fc2.weight"], sd[prefix + ".mlp.fc2.bias"])
    x = x + y
    return x.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()


def dframe_attention(sd, prefix, frames, depth, q_ind, num_heads, ws=(7, 7)):
    """DFrameAttention.forward (DTransformer.py:376-389): only the query frame is updated."""
    frames = list(frames)
    x = frames[q_ind]
    for i in range(depth):
        frames[q_ind] = x
        x = attention_block(sd, "%s.blocks.%d" % (prefix, i), frames, q_ind, num_heads, dilated=(i % 2 == 1), ws=ws)
    return x


# --------------------------------------------------------------------------------------
# BDE2VID generator forward  (bde2vid_cross_scale_propogation_V5.py:100-241)
# --------------------------------------------------------------------------------------

def _check_cfg(cfg):
    if cfg["norm"] not in (None, "none", "BN", "IN"):
        raise NotImplementedError("norm=%r" % (cfg["norm"],))
    if cfg["recurrent_block_type"] not in ("convlstm", "convgru"):
        raise NotImplementedError("recurrent_block_type=%r" % (cfg["recurrent_block_type"],))
    if cfg["skip_type"] not in ("sum", "concat"):
        raise NotImplementedError("skip_type=%r (no_skip crashes in the reference too)" % (cfg["skip_type"],))


def bde2vid_forward(sd, cfg, voxels, prefix="generator", taps=None):
    """voxels: list over time of [B, num_bins, Hp, Wp] fp32.  Returns list of [B, 1, Hp, Wp].

    ``taps`` (optional dict) receives intermediate tensors for stage-level parity tests.  The architecture variant
    (norm layers, ConvGRU, useRC=False, reduction conv, residual tail, concat skips) is read off the state_dict keys, as
    SURVEY.md appendix B describes; ``cfg`` supplies what has no parameters (buffer_index, q_idx, depths, heads, ...)."""
    cfg = full_cfg(cfg)
    _check_cfg(cfg)
    T = len(voxels)
    L = cfg["num_encoders"]
    buf = list(cfg["buffer_index"])
    q_ind = cfg["q_idx"]
    act = "ReLU" if cfg["act_net"] == "default" else cfg["act_net"]
    p = prefix + "." if prefix else ""
    use_rc = cfg.get("useRC", True)
    concat = cfg["skip_type"] == "concat"

    head = [conv_layer(sd, p + "head", v, 1, act) for v in voxels]                     # :116
    if taps is not None:
        taps["head"] = head
    x_seq = head
    levels = []
    for l in range(L):                                                                  # :119
        fwd, bwd = [None] * T, [None] * T
        sf = sb = None
        for k in range(T):                                                              # :122-135
            kb = T - 1 - k
            if use_rc:
                ef = conv_layer(sd, p + "forward_encoder.%d.conv" % l, x_seq[k], 2, act)
                fwd[k], sf = recurrent_step(sd, p + "forward_encoder.%d.recurrent_block" % l, ef, sf)
                eb = conv_layer(sd, p + "backward_encoder.%d.conv" % l, x_seq[kb], 2, act)
                bwd[kb], sb = recurrent_step(sd, p + "backward_encoder.%d.recurrent_block" % l, eb, sb)
            else:                                                                       # Encoder(), :255-257: ConvLayer only
                fwd[k] = conv_layer(sd, p + "forward_encoder.%d" % l, x_seq[k], 2, act)
                bwd[kb] = conv_layer(sd, p + "backward_encoder.%d" % l, x_seq[kb], 2, act)
        merged = [fwd[t] + bwd[t] for t in range(T)]                                    # :137-147
        if taps is not None:
            taps["merged%d" % l] = list(merged)
        depth = cfg["depths"][l]
        tail = l == L - 1 and depth == 0                                                # :77-80
        if depth > 0 or tail:                                                           # :151-169
            zero = torch.zeros_like(merged[0])
            for t in range(T):
                frames = [merged[t + o] if 0 <= t + o < T else zero for o in buf]       # Q1/Q4
                if tail:
                    # ParseLayer (x[0]) + ResidualBlockNoBN x num_res_blocks (:261-282), quirk Q5
                    x = frames[0]
                    for r in range(cfg["num_res_blocks"]):
                        rp = p + "feat_attns.%d.%d" % (l, r + 1)
                        y = _act(F.conv2d(x, sd[rp + ".conv1.weight"], sd[rp + ".conv1.bias"], padding=1), act)
                        x = x + F.conv2d(y, sd[rp + ".conv2.weight"], sd[rp + ".conv2.bias"], padding=1)
                else:
                    x = dframe_attention(sd, p + "feat_attns.%d" % l, frames, depth, q_ind, cfg["num_heads"],
                                         tuple(cfg["window_size"]))
                merged[t] = x + merged[t]
        if taps is not None:
            taps["level%d" % l] = list(merged)
        levels.append(merged)
        x_seq = merged

    skips = levels[:-1] + [levels[-1], levels[-1]]                                      # Q2 (:149-150,:172)
    out = []
    for t in range(T):                                                                  # :183-197
        x = skips[-1][t]
        for i in range(L):
            if concat:                                                                  # skip_concat + fusion conv (:86-93)
                x = torch.cat([skips[-2 - i][t], x], 1)
                x = F.conv2d(x, sd[p + "decoders.%d.0.weight" % i], sd[p + "decoders.%d.0.bias" % i])
            else:
                x = skips[-2 - i][t] + x
            x = upsample_conv(sd, p + "decoders.%d.1" % i, x, "ReLU6")
        if concat:
            x = F.conv2d(torch.cat([x, head[t]], 1), sd[p + "predI.0.weight"], sd[p + "predI.0.bias"])
        else:
            x = x + head[t]
        img = F.conv2d(x, sd[p + "predI.1.weight"], sd[p + "predI.1.bias"])
        if dict(cfg["activation"] or {}).get("type", "Sigmoid") == "Sigmoid":
            img = torch.sigmoid(img)
        out.append(img)
    return out


# --------------------------------------------------------------------------------------
# E2VIDRecurrent forward  (model/e2vid/unet.py:139-200; model/e2vid/submodules.py)
# --------------------------------------------------------------------------------------

def e2vid_recurrent_forward(sd, x, prev_states, num_encoders=4, num_residual_blocks=2, prefix="unetrecurrent"):
    """One step: x [B, bins, H, W], prev_states list of states (ConvLSTM (h, c) / ConvGRU h) or None -> (img, states)."""
    p = prefix + "." if prefix else ""
    x = conv_layer(sd, p + "head", x, 1, "relu")
    head = x
    if prev_states is None:
        prev_states = [None] * num_encoders
    blocks, states = [], []
    for i in range(num_encoders):
        x = conv_layer(sd, p + "encoders.%d.conv" % i, x, 2, "relu")
        x, st = recurrent_step(sd, p + "encoders.%d.recurrent_block" % i, x, prev_states[i])
        blocks.append(x)
        states.append(st)
    for r in range(num_residual_blocks):                       # ResidualBlock e2vid/submodules.py:212-247
        x = residual_block(sd, p + "resblocks.%d" % r, x)
    for i in range(num_encoders):
        x = upsample_conv(sd, p + "decoders.%d" % i, x + blocks[num_encoders - 1 - i], "relu")
    img = torch.sigmoid(F.conv2d(x + head, sd[p + "pred.conv2d.weight"], sd[p + "pred.conv2d.bias"]))
    return img, states


def residual_block(sd, prefix, x):
    """ResidualBlock with norm=None (model/submodules.py / model/e2vid/submodules.py:212-247): relu(conv2(relu(conv1 x)) + x)."""
    y = F.relu(F.conv2d(x, sd[prefix + ".conv1.weight"], sd[prefix + ".conv1.bias"], padding=1))
    y = F.conv2d(y, sd[prefix + ".conv2.weight"], sd[prefix + ".conv2.bias"], padding=1)
    return F.relu(y + x)


def firenet_forward(sd, x, states):
    """FireNet.forward (model/e2vid/model.py:119-172): head -> G1 (ConvGRU) -> R1 -> G2 -> R2 -> pred (1x1, no activation)."""
    if states is None:
        states = [None, None]
    x = conv_layer(sd, "head", x, 1, "relu")
    s1 = convgru_step(sd, "G1", x, states[0])
    x = residual_block(sd, "R1", s1)
    s2 = convgru_step(sd, "G2", x, states[1])
    x = residual_block(sd, "R2", s2)
    img = F.conv2d(x, sd["pred.conv2d.weight"], sd["pred.conv2d.bias"])
    return img, [s1, s2]


# --------------------------------------------------------------------------------------
# loader-side voxel transforms
# --------------------------------------------------------------------------------------

def loader_window(xs_i16, ys_i16, ts_f64, ps_bool):
    """DynamicH5Dataset.get_events + __getitem__ conversions (h5_dataset.py:410-415, :222-225) of one window given in the
    on-disk dtypes: ps * 2.0 - 1.0, float32 casts, ts - ts[0] in float64 before the cast."""
    ps = ps_bool * 2.0 - 1.0
    ts0 = ts_f64[0]
    return (xs_i16.astype(np.float32), ys_i16.astype(np.float32), (ts_f64 - ts0).astype(np.float32), ps.astype(np.float32))


def loader_voxel(xs_i16, ys_i16, ts_f64, ps_bool, num_bins, sensor_size, hot_mask=None):
    """__getitem__'s voxel (h5_dataset.py:219-226, :357-364): fewer than 3 events -> zeros; else the voxel grid times the
    hot-pixel mask."""
    if len(xs_i16) < 3:
        return np.zeros((num_bins,) + tuple(sensor_size), dtype=np.float32)
    g = voxel_grid(*loader_window(xs_i16, ys_i16, ts_f64, ps_bool), num_bins, sensor_size)
    if hot_mask is not None:
        g = g * hot_mask[None].astype(np.float32)
    return g


def legacy_norm(x):
    """LegacyNorm.__call__ (utils_func/data_augmentation.py:316-330) on a float32 tensor."""
    x = torch.as_tensor(x)
    nonzero = (x != 0)
    num_nonzeros = nonzero.sum()
    if num_nonzeros > 0:
        mean = x.sum() / num_nonzeros
        stddev = torch.sqrt((x ** 2).sum() / num_nonzeros - mean ** 2)
        if stddev != 0:
            mask = nonzero.float()
            x = mask * (x - mean) / stddev
    return x


def robust_norm(x, low_perc=0, top_perc=95):
    """RobustNorm.__call__ (utils_func/utils.py:17-51): nearest-rank percentiles via kthvalue, clamp, rescale."""
    x = torch.as_tensor(x)

    def percentile(t, q):
        k = 1 + round(.01 * float(q) * (t.numel() - 1))
        return t.reshape(-1).kthvalue(k).values.item()

    t_max = percentile(x, top_perc)
    t_min = percentile(x, low_perc)
    if t_max == 0 and t_min == 0:
        return x
    normed = torch.clamp(x, min=t_min, max=t_max)
    return (normed - torch.min(normed)) / (torch.max(normed) + 1e-6)


def hot_event_mask(xs, ys, ps, sensor_size, num_hot):
    """get_hot_event_mask (event_utils.py:100-116) with events_to_image's bincount accumulation (:155-174); ps = +-1."""
    H, W = sensor_size
    img = np.bincount(np.asarray(ys, dtype=np.int64) * W + np.asarray(xs, dtype=np.int64), weights=np.asarray(ps, dtype=np.float64),
                      minlength=H * W).reshape(H, W)
    mask = np.ones_like(img)
    for _ in range(num_hot):
        idx = np.unravel_index(np.argmax(img), img.shape)
        mask[idx] = 0
        img[idx] = 0
    return mask


# --------------------------------------------------------------------------------------
# metrics used for the frame gates (evaluate/metrics.py:42-65)
# --------------------------------------------------------------------------------------

def mse(a, b):
    return float(F.mse_loss(a, b))


def ssim_uniform7(a, b, data_range=1.0):
    """skimage.metrics.structural_similarity defaults restated (7x7 uniform window, K1=.01,
    K2=.03, sample covariance, mean over the valid interior).  skimage is absent here, so this
    metric is 'parity unpinned' (SURVEY.md 8(c)); it is only used for our-vs-oracle deltas."""
    a = a.reshape(1, 1, *a.shape[-2:]).double()
    b = b.reshape(1, 1, *b.shape[-2:]).double()
    win = 7
    NP = win * win
    cov_norm = NP / (NP - 1.0)
    k = torch.ones(1, 1, win, win, dtype=torch.float64) / NP
    ux, uy = F.conv2d(a, k), F.conv2d(b, k)
    uxx, uyy, uxy = F.conv2d(a * a, k), F.conv2d(b * b, k), F.conv2d(a * b, k)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    return float(S.mean())
