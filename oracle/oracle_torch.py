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


def conv_layer(sd, prefix, x, stride, act):
    """ConvLayer (submodules.py:85-114), norm=None."""
    w = sd[prefix + ".conv2d.weight"]
    return _act(F.conv2d(x, w, sd[prefix + ".conv2d.bias"], stride=stride, padding=w.shape[-1] // 2), act)


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


def upsample_conv(sd, prefix, x, act):
    """UpsampleConvLayer (submodules.py:137-147): bilinear x2 (align_corners=False) then conv."""
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    w = sd[prefix + ".conv2d.weight"]
    return _act(F.conv2d(x, w, sd[prefix + ".conv2d.bias"], padding=w.shape[-1] // 2), act)


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


def rel_pos_bias(sd, prefix, D, wh, ww, q_ind, num_heads):
    """relative_position_bias_table gathered for the query frame's tokens
    (DTransformer.py:139-152, 195-199) -> [nH, wh*ww, D*wh*ww]."""
    table = sd[prefix + ".relative_position_bias_table"]
    index = sd[prefix + ".relative_position_index"]
    n = wh * ww
    idx = index[q_ind * n:(q_ind + 1) * n, :D * n].reshape(-1)
    return table[idx].reshape(n, D * n, num_heads).permute(2, 0, 1).contiguous()


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
    kv_in = torch.cat(tok, dim=2)                                   # [B, nWin, D*n, C], index d*n + a*ww + b
    qn = F.layer_norm(q_in, (C,), sd[ap + ".norm_q.weight"], sd[ap + ".norm_q.bias"], 1e-5)
    kvn = F.layer_norm(kv_in, (C,), sd[ap + ".norm_kv.weight"], sd[ap + ".norm_kv.bias"], 1e-5)
    q = F.linear(qn, sd[ap + ".q.weight"], sd[ap + ".q.bias"])
    kv = F.linear(kvn, sd[ap + ".kv.weight"], sd[ap + ".kv.bias"])
    q = q.reshape(B, nWin, n, num_heads, hd).permute(0, 1, 3, 2, 4) * (hd ** -0.5)
    k = kv[..., :C].reshape(B, nWin, D * n, num_heads, hd).permute(0, 1, 3, 2, 4)
    v = kv[..., C:].reshape(B, nWin, D * n, num_heads, hd).permute(0, 1, 3, 2, 4)
    attn = q @ k.transpose(-2, -1)
    attn = attn + rel_pos_bias(sd, ap, D, g["wh"], g["ww"], q_ind, num_heads).view(1, 1, num_heads, n, D * n)
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
    if cfg["norm"] not in (None, "none"):
        raise NotImplementedError("norm=%r" % (cfg["norm"],))
    if cfg["recurrent_block_type"] != "convlstm" or not cfg["useRC"]:
        raise NotImplementedError("only useRC=True with convlstm is restated")
    if cfg["skip_type"] != "sum":
        raise NotImplementedError("skip_type=%r" % (cfg["skip_type"],))
    if cfg["nwindow_size"] is not None:
        raise NotImplementedError("nwindow_size")
    if cfg["depths"][-1] == 0:
        raise NotImplementedError("last-level depth 0 (ParseLayer path)")


def bde2vid_forward(sd, cfg, voxels, prefix="generator", taps=None):
    """voxels: list over time of [B, num_bins, Hp, Wp] fp32.  Returns list of [B, 1, Hp, Wp].

    ``taps`` (optional dict) receives intermediate tensors for stage-level parity tests."""
    cfg = full_cfg(cfg)
    _check_cfg(cfg)
    T = len(voxels)
    L = cfg["num_encoders"]
    buf = list(cfg["buffer_index"])
    q_ind = cfg["q_idx"]
    act = "ReLU" if cfg["act_net"] == "default" else cfg["act_net"]
    p = prefix + "." if prefix else ""

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
            ef = conv_layer(sd, p + "forward_encoder.%d.conv" % l, x_seq[k], 2, act)
            sf = convlstm_step(sd, p + "forward_encoder.%d.recurrent_block" % l, ef, sf)
            fwd[k] = sf[0]
            eb = conv_layer(sd, p + "backward_encoder.%d.conv" % l, x_seq[kb], 2, act)
            sb = convlstm_step(sd, p + "backward_encoder.%d.recurrent_block" % l, eb, sb)
            bwd[kb] = sb[0]
        merged = [fwd[t] + bwd[t] for t in range(T)]                                    # :137-147
        if taps is not None:
            taps["merged%d" % l] = list(merged)
        depth = cfg["depths"][l]
        if depth > 0:                                                                   # :151-169
            zero = torch.zeros_like(merged[0])
            for t in range(T):
                frames = [merged[t + o] if 0 <= t + o < T else zero for o in buf]       # Q1/Q4
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
            x = upsample_conv(sd, p + "decoders.%d.1" % i, skips[-2 - i][t] + x, "ReLU6")
        x = x + head[t]
        img = torch.sigmoid(F.conv2d(x, sd[p + "predI.1.weight"], sd[p + "predI.1.bias"]))
        out.append(img)
    return out


# --------------------------------------------------------------------------------------
# E2VIDRecurrent forward  (model/e2vid/unet.py:139-200; model/e2vid/submodules.py)
# --------------------------------------------------------------------------------------

def e2vid_recurrent_forward(sd, x, prev_states, num_encoders=4, num_residual_blocks=2, prefix="unetrecurrent"):
    """One step: x [B, bins, H, W], prev_states list of (h, c) or None -> (img, states)."""
    p = prefix + "." if prefix else ""
    x = conv_layer(sd, p + "head", x, 1, "relu")
    head = x
    if prev_states is None:
        prev_states = [None] * num_encoders
    blocks, states = [], []
    for i in range(num_encoders):
        x = conv_layer(sd, p + "encoders.%d.conv" % i, x, 2, "relu")
        st = convlstm_step(sd, p + "encoders.%d.recurrent_block" % i, x, prev_states[i])
        x = st[0]
        blocks.append(x)
        states.append(st)
    for r in range(num_residual_blocks):                       # ResidualBlock e2vid/submodules.py:212-247
        rp = p + "resblocks.%d" % r
        y = F.relu(F.conv2d(x, sd[rp + ".conv1.weight"], sd[rp + ".conv1.bias"], padding=1))
        y = F.conv2d(y, sd[rp + ".conv2.weight"], sd[rp + ".conv2.bias"], padding=1)
        x = F.relu(y + x)
    for i in range(num_encoders):
        x = upsample_conv(sd, p + "decoders.%d" % i, x + blocks[num_encoders - 1 - i], "relu")
    img = torch.sigmoid(F.conv2d(x + head, sd[p + "pred.conv2d.weight"], sd[p + "pred.conv2d.bias"]))
    return img, states


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
