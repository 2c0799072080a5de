"""Synthetic event streams and synthetic checkpoints.

``gen_events`` produces events in the reference's on-disk dtypes
(events_contrast_maximization/tools/event_packagers.py:44-47: xs,ys int16, ts float64 seconds,
ps bool) and ``to_loader_format`` converts one window exactly as the loader does
(data_loader/h5_dataset.py:222-225, :414): x,y -> float32, (t - t[0]) -> float32, p -> +-1 float32.
"""
import numpy as np
import torch

WINDOW_DT = 1.0 / 30.0


def gen_events(seq_id, T, H, W, N):
    """T windows of N events each.  Returns dict of numpy arrays + CSR ``offsets`` [T+1]."""
    g = torch.Generator().manual_seed(1000 + int(seq_id))
    xs = torch.randint(0, W, (T, N), generator=g, dtype=torch.int16)
    ys = torch.randint(0, H, (T, N), generator=g, dtype=torch.int16)
    ps = torch.rand(T, N, generator=g) < 0.5
    u, _ = torch.sort(torch.rand(T, N, generator=g, dtype=torch.float64), dim=1)
    ts = u * WINDOW_DT + torch.arange(T, dtype=torch.float64).view(T, 1) * WINDOW_DT
    return dict(xs=xs.reshape(-1).numpy(), ys=ys.reshape(-1).numpy(), ts=ts.reshape(-1).numpy(),
                ps=ps.reshape(-1).numpy(), offsets=np.arange(T + 1, dtype=np.int64) * N, H=H, W=W)


def to_loader_format(ev, w):
    """Window ``w`` as four float32 arrays, the voxeliser's input contract."""
    a, b = int(ev["offsets"][w]), int(ev["offsets"][w + 1])
    xs = ev["xs"][a:b].astype(np.float32)
    ys = ev["ys"][a:b].astype(np.float32)
    ts = ev["ts"][a:b]
    ts = (ts - ts[0]).astype(np.float32) if b > a else ts.astype(np.float32)    # (empty window: nothing to shift)
    ps = ev["ps"][a:b].astype(np.float32) * 2.0 - 1.0
    return xs, ys, ts, ps


def to_loader_format_seq(ev):
    """All windows concatenated (per-window relative fp32 timestamps) + offsets."""
    T = len(ev["offsets"]) - 1
    parts = [to_loader_format(ev, w) for w in range(T)]
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(4)) + (ev["offsets"].copy(),)


ASSUMED_CFG_STR = """model = dict(
    type='BDE2VID',
    generator=dict(
        type='BDE2VIDCrossscalePropogationV5',
        num_bins=5, basechannels=32, num_encoders=3, ks=5, num_res_blocks=2, norm=None,
        recurrent_block_type='convlstm', useRC=True, skip_type='sum',
        buffer_index=[-1, 0, 1], q_idx=1, window_size=(7, 7), nwindow_size=None,
        depths=[4, 0, 6], num_heads=16, losses=[]))
"""


# --------------------------------------------------------------------------------------
# synthetic checkpoints
# --------------------------------------------------------------------------------------

def state_dict_spec(cfg):
    """Ordered {key: (shape, kind)} of the reference's BDE2VID state_dict for a generator cfg
    (model/BDE2VID/bde2vid_cross_scale_propogation_V5.py:19-98).
    kind in {'w', 'b', 'ln_w', 'ln_b', 'table', 'index'}."""
    nb, bc, ne, ks = cfg["num_bins"], cfg["basechannels"], cfg["num_encoders"], cfg["ks"]
    depths = list(cfg.get("depths", [4, 0, 6]))
    heads = cfg.get("num_heads", 16)
    wh, ww = tuple(cfg.get("window_size", (7, 7)))
    D = len(cfg["buffer_index"])
    spec = {}

    def conv(name, co, ci, k):
        spec[name + ".weight"] = ((co, ci, k, k), "w")
        spec[name + ".bias"] = ((co,), "b")

    def lin(name, co, ci):
        spec[name + ".weight"] = ((co, ci), "w")
        spec[name + ".bias"] = ((co,), "b")

    def ln(name, c):
        spec[name + ".weight"] = ((c,), "ln_w")
        spec[name + ".bias"] = ((c,), "ln_b")

    g = "generator."
    conv(g + "head.conv2d", bc, nb, ks)
    for enc in ("forward_encoder", "backward_encoder"):
        for i in range(ne):
            ci, co = bc * 2 ** i, bc * 2 ** (i + 1)
            conv(g + "%s.%d.conv.conv2d" % (enc, i), co, ci, ks)
            conv(g + "%s.%d.recurrent_block.Gates" % (enc, i), 4 * co, 2 * co, 3)
    for i in range(ne):
        co = bc * 2 ** (i + 1)
        conv(g + "fusion_layers.%d" % i, co, 2 * co, 1)
    for i, depth in enumerate(depths):
        c = bc * 2 ** (i + 1)
        for b in range(depth):
            p = g + "feat_attns.%d.blocks.%d." % (i, b)
            spec[p + "attn.relative_position_bias_table"] = (((2 * D - 1) * (2 * wh - 1) * (2 * ww - 1), heads), "table")
            spec[p + "attn.relative_position_index"] = ((D * wh * ww, D * wh * ww), "index")
            ln(p + "attn.norm_q", c)
            ln(p + "attn.norm_kv", c)
            lin(p + "attn.q", c, c)
            lin(p + "attn.kv", 2 * c, c)
            lin(p + "attn.proj", c, c)
            ln(p + "norm2", c)
            lin(p + "mlp.fc1", 4 * c, c)
            lin(p + "mlp.fc2", c, 4 * c)
    for i in range(ne):
        ci, co = bc * 2 ** (ne - i), bc * 2 ** (ne - i - 1)
        conv(g + "decoders.%d.1.conv2d" % i, co, ci, ks)
    conv(g + "predI.1", cfg.get("num_output_channels", 1), bc, 1)
    return spec


def relative_position_index(D, wh, ww):
    """idx[p, q] = ((dp-dq)+D-1)*(2wh-1)*(2ww-1) + ((ap-aq)+wh-1)*(2ww-1) + (bp-bq)+ww-1
    (model/BDE2VID/DTransformer.py:139-152)."""
    d, a, b = torch.meshgrid(torch.arange(D), torch.arange(wh), torch.arange(ww), indexing="ij")
    d, a, b = d.reshape(-1), a.reshape(-1), b.reshape(-1)
    return ((d[:, None] - d[None, :] + D - 1) * ((2 * wh - 1) * (2 * ww - 1))
            + (a[:, None] - a[None, :] + wh - 1) * (2 * ww - 1)
            + (b[:, None] - b[None, :] + ww - 1)).to(torch.int64)


def init_state_dict(cfg, seed=0, stress=False):
    """Seeded random checkpoint with the reference's key set.  Weights/biases are
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (PyTorch's default conv/linear init, which the reference
    uses), LayerNorm = (1, 0), bias table ~ N(0, .02).  ``stress=True`` perturbs LayerNorm affine
    parameters and widens the bias table so that zero-token / bias paths are exercised by tests."""
    g = torch.Generator().manual_seed(int(seed))
    D = len(cfg["buffer_index"])
    wh, ww = tuple(cfg.get("window_size", (7, 7)))
    sd = {}
    spec = state_dict_spec(cfg)
    for key, (shape, kind) in spec.items():
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / fan_in ** 0.5
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "b":
            wshape = spec[key[:-4] + "weight"][0]
            fan_in = 1
            for s in wshape[1:]:
                fan_in *= s
            bound = 1.0 / fan_in ** 0.5
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "ln_w":
            sd[key] = torch.ones(shape) + (0.2 * torch.randn(shape, generator=g) if stress else 0)
        elif kind == "ln_b":
            sd[key] = 0.2 * torch.randn(shape, generator=g) if stress else torch.zeros(shape)
        elif kind == "table":
            sd[key] = (0.5 if stress else 0.02) * torch.randn(shape, generator=g).clamp_(-2, 2)
        elif kind == "index":
            sd[key] = relative_position_index(D, wh, ww)
    return sd


def random_state_dict_like(template, seed=0):
    """Seeded random checkpoint with the key set / shapes of ``template`` (a state_dict of the reference model or of the
    drop-in container -- they have the same keys).  Every tensor is drawn from a generator seeded by (crc32(key), seed),
    so the result does not depend on key order.  Used for the architecture variants (ConvGRU, BN / IN, concat skips,
    nwindow_size, residual tail, FireNet ...) whose key sets differ from ``state_dict_spec``."""
    import zlib
    out = {}
    for key, v in template.items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * int(seed)) & 0x7fffffff)
        shape = tuple(v.shape)
        if not torch.is_floating_point(v):
            out[key] = v.clone()                                   # relative_position_index, num_batches_tracked
        elif key.endswith("running_var"):
            out[key] = 0.5 + torch.rand(shape, generator=g)
        elif key.endswith("running_mean"):
            out[key] = 0.2 * torch.randn(shape, generator=g)
        elif key.endswith("relative_position_bias_table"):
            out[key] = 0.5 * torch.randn(shape, generator=g).clamp_(-2, 2)
        elif ("norm" in key.rsplit(".", 2)[-2]) and key.endswith(".weight"):
            out[key] = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif ("norm" in key.rsplit(".", 2)[-2]) and key.endswith(".bias"):
            out[key] = 0.2 * torch.randn(shape, generator=g)
        elif v.dim() >= 2:
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            out[key] = (torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5
        else:
            out[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
    return out


def gen_events_clustered(seq_id, T, H, W, N, blobs=24, sigma=6.0):
    """Like ``gen_events`` but spatially clustered: every window's events fall around ``blobs`` moving Gaussian blobs
    (sigma pixels), the way real event streams concentrate on moving edges.  Same dtypes / CSR layout."""
    g = torch.Generator().manual_seed(5000 + int(seq_id))
    cx = torch.rand(blobs, generator=g) * W
    cy = torch.rand(blobs, generator=g) * H
    vx = (torch.rand(blobs, generator=g) - 0.5) * 8
    vy = (torch.rand(blobs, generator=g) - 0.5) * 8
    xs, ys = [], []
    for t in range(T):
        which = torch.randint(0, blobs, (N,), generator=g)
        x = (cx + vx * t)[which] + sigma * torch.randn(N, generator=g)
        y = (cy + vy * t)[which] + sigma * torch.randn(N, generator=g)
        xs.append(x.remainder(W).floor().clamp_(0, W - 1).to(torch.int16))
        ys.append(y.remainder(H).floor().clamp_(0, H - 1).to(torch.int16))
    ps = torch.rand(T, N, generator=g) < 0.5
    u, _ = torch.sort(torch.rand(T, N, generator=g, dtype=torch.float64), dim=1)
    ts = u * WINDOW_DT + torch.arange(T, dtype=torch.float64).view(T, 1) * WINDOW_DT
    return dict(xs=torch.cat(xs).numpy(), ys=torch.cat(ys).numpy(), ts=ts.reshape(-1).numpy(), ps=ps.reshape(-1).numpy(),
                offsets=np.arange(T + 1, dtype=np.int64) * N, H=H, W=W)
