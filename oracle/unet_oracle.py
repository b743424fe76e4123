"""CPU oracle for the U-Net diffusion training step of clu0/unet.cu  --  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch (CPU, fp32 or fp64), the algorithm of the reference's hot
path.  It is the checker the CUDA path is compared with; nothing in the product (unet.cu_b200/) imports it.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.

Parity pinning: tests/test_oracle_vs_reference.py runs this restatement against the reference's own Python
modules (/root/reference/dev/unet.py, dev/resblock.py, train_unet.py) when the reference tree is present
(i.e. in the build container), and tests/test_oracle_golden.py pins it against the committed fixtures in
tests/golden/ that oracle/gen_golden.py produced from the reference itself.

Each function cites the reference lines it follows (paths under /root/reference).

One restated feature has no reference implementation to pin against and says so: dropout (UNetConfig.dropout) is
commented out in the reference (dev/resblock.py:51,61, dev/unet.py:116,140) -- PARITY UNPINNED for it; the restatement is
nn.Dropout's training-mode arithmetic with a given keep-mask.  Everything else is pinned (class labels, resblock_updown:
the reference's UNetModel; use_scale_shift_norm: the reference's ResBlockO).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MODEL_MAGIC = 12345678  # train_unet.py:781, train_unet.cu:4826
DATA_MAGIC = 20240620   # prepare_data.py:20-38


# ----------------------------------------------------------------------------------------------- config
@dataclass(frozen=True)
class UNetConfig:
    """Mirror of UnetConfig (train_unet.cu:3318-3336) / UNetModel ctor args (dev/unet.py:129-145)."""
    in_channels: int = 3
    model_channels: int = 64
    out_channels: int = 3
    num_res_blocks: int = 2
    channel_mult: Tuple[int, ...] = (1, 2, 3, 4)
    attn_start_level: int = 2          # attention at levels >= this (train_unet.cu:4858: att_start_level = 2)
    head_size: int = 32                # num_head_channels (dev/unet_test.py:253)
    gn_groups: int = 32
    max_period: int = 1000             # header[7], dev/unet.py:327
    H: int = 64
    W: int = 64
    num_classes: int = 0               # > 0: class-conditional, label_emb = nn.Embedding(num_classes, 4*mc) (dev/unet.py:174-175)
    resblock_updown: bool = False      # ResBlock(down=True / up=True) instead of Downsample / Upsample (dev/unet.py:147,205-222,271-284)
    use_scale_shift_norm: bool = False # FiLM-like conditioning in every ResBlock (dev/unet.py:146, dev/resblock.py:211,243-247)
    dropout: float = 0.0               # Dropout(p) between SiLU and conv2 of every ResBlock (commented out in the reference:
                                       # dev/resblock.py:51,61; guided-diffusion's out_layers = GroupNorm, SiLU, Dropout, conv)

    @property
    def emb_channels(self) -> int:
        return 4 * self.model_channels

    def attention_resolutions(self) -> Tuple[int, ...]:
        """Downsample rates `ds` at which dev/unet.py places attention (dev/unet.py:196, 262)."""
        return tuple(2 ** l for l in range(self.attn_start_level, len(self.channel_mult)))


# One entry per layer in forward order; this IS the parameter / .bin order (train_unet.cu:4335-4421,
# dev/unet.py:183-291: time_embed, input_blocks, middle_block, output_blocks, out).
@dataclass
class Layer:
    kind: str                 # 'linear' 'conv3' 'res' 'attn' 'down' 'up' 'gn' 'concat' 'silu'
    cin: int = 0
    cout: int = 0
    level: int = 0
    params: List[Tuple[str, Tuple[int, ...]]] = field(default_factory=list)
    skip_push: bool = False   # output is pushed on the skip stack (dev/unet.py:314)
    name: str = ""


def res_params(prefix: str, cin: int, cout: int, cemb: int, scale_shift: bool = False):
    """ResBlock parameter order (dev/resblock.cuh:8-21, dev/resblock.py:70-105); with use_scale_shift_norm the embedding
    projection has 2 * cout outputs (dev/resblock.py:91-94)."""
    oce = 2 * cout if scale_shift else cout
    p = [(f"{prefix}.gn1.weight", (cin,)), (f"{prefix}.gn1.bias", (cin,)),
         (f"{prefix}.cv3_1.weight", (cout, cin, 3, 3)), (f"{prefix}.cv3_1.bias", (cout,)),
         (f"{prefix}.l_emb.weight", (oce, cemb)), (f"{prefix}.l_emb.bias", (oce,)),
         (f"{prefix}.gn2.weight", (cout,)), (f"{prefix}.gn2.bias", (cout,)),
         (f"{prefix}.cv3_2.weight", (cout, cout, 3, 3)), (f"{prefix}.cv3_2.bias", (cout,))]
    if cin != cout:
        p += [(f"{prefix}.skip_connection.weight", (cout, cin, 1, 1)), (f"{prefix}.skip_connection.bias", (cout,))]
    return p


def attn_params(prefix: str, c: int):
    """AttentionBlock parameter order (dev/attention_block.cuh:4-12, dev/unet.py:37-42)."""
    return [(f"{prefix}.gn.weight", (c,)), (f"{prefix}.gn.bias", (c,)),
            (f"{prefix}.qkv.weight", (3 * c, c, 1)), (f"{prefix}.qkv.bias", (3 * c,)),
            (f"{prefix}.proj.weight", (c, c, 1)), (f"{prefix}.proj.bias", (c,))]


def build_layers(cfg: UNetConfig) -> List[Layer]:
    """Layer list in forward order; follows dev/unet.py:165-291 and train_unet.cu:3447-3487."""
    mc, cemb = cfg.model_channels, cfg.emb_channels
    L: List[Layer] = []
    L.append(Layer('linear', mc, cemb, params=[("time_embed.0.weight", (cemb, mc)), ("time_embed.0.bias", (cemb,))],
                   name="time_embed.0"))
    L.append(Layer('linear', cemb, cemb, params=[("time_embed.2.weight", (cemb, cemb)),
                                                   ("time_embed.2.bias", (cemb,))], name="time_embed.2"))
    if cfg.num_classes:   # dev/unet.py:174-175: constructed (and hence ordered) right after time_embed
        L.append(Layer('embed', cfg.num_classes, cemb, params=[("label_emb.weight", (cfg.num_classes, cemb))],
                       name="label_emb"))
    ch = cfg.channel_mult[0] * mc
    L.append(Layer('conv3', cfg.in_channels, ch, params=[("input_blocks.0.0.weight", (ch, cfg.in_channels, 3, 3)),
                                                          ("input_blocks.0.0.bias", (ch,))], skip_push=True,
                   name="input_blocks.0.0"))
    chans = [ch]
    ib = 1
    nlev = len(cfg.channel_mult)
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            cout = mult * mc
            has_attn = level >= cfg.attn_start_level
            L.append(Layer('res', ch, cout, level, res_params(f"input_blocks.{ib}.0", ch, cout, cemb, cfg.use_scale_shift_norm),
                           skip_push=not has_attn, name=f"input_blocks.{ib}.0"))
            ch = cout
            if has_attn:
                L.append(Layer('attn', ch, ch, level, attn_params(f"input_blocks.{ib}.1", ch), skip_push=True,
                               name=f"input_blocks.{ib}.1"))
            chans.append(ch)
            ib += 1
        if level != nlev - 1:
            if cfg.resblock_updown:
                L.append(Layer('res_down', ch, ch, level, res_params(f"input_blocks.{ib}.0", ch, ch, cemb, cfg.use_scale_shift_norm), skip_push=True,
                               name=f"input_blocks.{ib}.0"))
            else:
                L.append(Layer('down', ch, ch, level, skip_push=True, name=f"input_blocks.{ib}.0"))
            chans.append(ch)
            ib += 1
    L.append(Layer('res', ch, ch, nlev - 1, res_params("middle_block.0", ch, ch, cemb, cfg.use_scale_shift_norm), name="middle_block.0"))
    L.append(Layer('attn', ch, ch, nlev - 1, attn_params("middle_block.1", ch), name="middle_block.1"))
    L.append(Layer('res', ch, ch, nlev - 1, res_params("middle_block.2", ch, ch, cemb, cfg.use_scale_shift_norm), name="middle_block.2"))
    ob = 0
    for level in reversed(range(nlev)):
        mult = cfg.channel_mult[level]
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            cout = mult * mc
            L.append(Layer('concat', ch, ch + ich, level, name=f"output_blocks.{ob}.cat"))
            L.append(Layer('res', ch + ich, cout, level, res_params(f"output_blocks.{ob}.0", ch + ich, cout, cemb, cfg.use_scale_shift_norm),
                           name=f"output_blocks.{ob}.0"))
            ch = cout
            sub = 1
            if level >= cfg.attn_start_level:
                L.append(Layer('attn', ch, ch, level, attn_params(f"output_blocks.{ob}.1", ch),
                               name=f"output_blocks.{ob}.1"))
                sub = 2
            if level and i == cfg.num_res_blocks:
                if cfg.resblock_updown:
                    L.append(Layer('res_up', ch, ch, level, res_params(f"output_blocks.{ob}.{sub}", ch, ch, cemb, cfg.use_scale_shift_norm),
                                   name=f"output_blocks.{ob}.{sub}"))
                else:
                    L.append(Layer('up', ch, ch, level, name=f"output_blocks.{ob}.{sub}"))
            ob += 1
    L.append(Layer('gn', ch, ch, 0, [("out.0.weight", (ch,)), ("out.0.bias", (ch,))], name="out.0"))
    L.append(Layer('conv3', ch, cfg.out_channels, 0, [("out.2.weight", (cfg.out_channels, ch, 3, 3)),
                                                       ("out.2.bias", (cfg.out_channels,))], name="out.2"))
    return L


def param_spec(cfg: UNetConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    return [p for l in build_layers(cfg) for p in l.params]


def num_params(cfg: UNetConfig) -> int:
    return sum(int(np.prod(s)) for _, s in param_spec(cfg))


def init_params(cfg: UNetConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """PyTorch default initialisation drawn in construction order, so that `torch.manual_seed(seed)` followed by
    this function yields the same tensors as `torch.manual_seed(seed); UNetModel(...)` (dev/unet.py:165-291;
    modules are constructed in parameter order, GroupNorm consumes no random numbers)."""
    torch.manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    spec = param_spec(cfg)
    i = 0
    while i < len(spec):
        name, shape = spec[i]
        base = name.rsplit('.', 1)[0]
        if name == 'label_emb.weight':                         # nn.Embedding: N(0, 1), no bias
            out[name] = torch.nn.Embedding(shape[0], shape[1]).weight.detach().to(dtype).clone()
            i += 1
            continue
        if len(shape) == 1 and name.endswith('.weight'):      # GroupNorm affine: ones / zeros
            out[name] = torch.ones(shape, dtype=dtype)
            out[spec[i + 1][0]] = torch.zeros(spec[i + 1][1], dtype=dtype)
        else:
            if len(shape) == 4:
                m = torch.nn.Conv2d(shape[1], shape[0], shape[2], padding=shape[2] // 2)
            elif len(shape) == 3:
                m = torch.nn.Conv1d(shape[1], shape[0], 1)
            else:
                m = torch.nn.Linear(shape[1], shape[0])
            out[base + '.weight'] = m.weight.detach().to(dtype).clone()
            out[base + '.bias'] = m.bias.detach().to(dtype).clone()
        i += 2
    return out


def flatten_params(cfg: UNetConfig, params: Dict[str, torch.Tensor]) -> torch.Tensor:
    return torch.cat([params[n].reshape(-1) for n, _ in param_spec(cfg)])


def unflatten_params(cfg: UNetConfig, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    for n, s in param_spec(cfg):
        k = int(np.prod(s))
        out[n] = flat[off:off + k].reshape(s)
        off += k
    assert off == flat.numel()
    return out


# ----------------------------------------------------------------------------------------------- .bin I/O
def write_model_bin(path: str, cfg: UNetConfig, flat: np.ndarray, B: int = 32, m: np.ndarray | None = None,
                    v: np.ndarray | None = None):
    """train_unet.py:768-795 (writer) / train_unet.cu:4762-4814: int32[256] header then float32 params
    [+ m + v when header[8]==1]; rng blob never written (header[9]=0)."""
    header = np.zeros(256, dtype=np.int32)
    header[:10] = [MODEL_MAGIC, B, cfg.in_channels, cfg.model_channels, cfg.out_channels, cfg.H, cfg.W,
                   cfg.max_period, 1 if m is not None else 0, 0]
    with open(path, 'wb') as f:
        f.write(header.tobytes())
        f.write(np.ascontiguousarray(flat, dtype=np.float32).tobytes())
        if m is not None:
            f.write(np.ascontiguousarray(m, dtype=np.float32).tobytes())
            f.write(np.ascontiguousarray(v, dtype=np.float32).tobytes())


def read_model_bin(path: str):
    """generate.py:17-27 / train_unet.cu:4819-4911."""
    with open(path, 'rb') as f:
        header = np.frombuffer(f.read(1024), dtype=np.int32)
        assert header[0] == MODEL_MAGIC, "bad magic"
        rest = np.frombuffer(f.read(), dtype=np.float32)
    return header.copy(), rest.copy()


def write_data_bin(path: str, images: np.ndarray):
    """prepare_data.py:20-38: int32[256] header {magic, N, C, H, W} + float32 images in [-1, 1]."""
    n, c, h, w = images.shape
    header = np.zeros(256, dtype=np.int32)
    header[:5] = [DATA_MAGIC, n, c, h, w]
    with open(path, 'wb') as f:
        f.write(header.tobytes())
        f.write(np.ascontiguousarray(images, dtype=np.float32).tobytes())


# ----------------------------------------------------------------------------------------------- layers
def timestep_embedding(t: torch.Tensor, dim: int, max_period: int = 1000) -> torch.Tensor:
    """dev/unet.py:327-345, train_unet.cu:3258-3313: [cos(t f_j), sin(t f_j)], f_j = exp(-ln(P) j / half).
    `t` has shape (B,) or (B,1)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.dtype)
    args = t.reshape(-1, 1) * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def silu(x):
    """train_unet.cu:316: x / (1 + exp(-x))."""
    return x * torch.sigmoid(x)


def groupnorm(x, w, b, groups=32, eps=1e-5):
    """train_unet.cu:1768-1828 (biased variance over each (sample, group)), eps 1e-5."""
    B, C = x.shape[:2]
    xg = x.reshape(B, groups, -1)
    mean = xg.mean(dim=2, keepdim=True)
    var = xg.var(dim=2, unbiased=False, keepdim=True)
    y = ((xg - mean) * torch.rsqrt(var + eps)).reshape(x.shape)
    shape = (1, C) + (1,) * (x.dim() - 2)
    return y * w.reshape(shape) + b.reshape(shape)


def conv3x3(x, w, b):
    """dev/conv2d_k3.py:15-114 (9-tap shift algebra) == cross-correlation, pad 1, stride 1."""
    return F.conv2d(x, w, b, padding=1)


def conv1x1(x, w, b):
    """dev/conv2d_k1.py:9-62."""
    return F.conv2d(x, w.reshape(w.shape[0], w.shape[1], 1, 1), b)


def avgpool2(x):
    """dev/downsample.py:10-32 / train_unet.cu:455-533."""
    return F.avg_pool2d(x, 2, 2)


def upsample2(x):
    """dev/upsample.py:8-27 / train_unet.cu:360-440: nearest x2."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def resblock(x, emb, P: Dict[str, torch.Tensor], prefix: str, groups=32, updown: str | None = None,
             drop_mask: torch.Tensor | None = None, drop_p: float = 0.0):
    """dev/resblock.py:107-160, train_unet.cu:2213-2287 (up = down = 0 in the default U-Net, train_unet.cu:4278).
    updown = 'down' / 'up' (dev/resblock.py:78-86, 125-128): h_upd / x_upd resample the main path between SiLU and conv1
    and the skip path -- average pool 2x2 (Downsample, dev/resblock.py:34-43) or nearest x2 (Upsample, :25-32)."""
    h = silu(groupnorm(x, P[prefix + '.gn1.weight'], P[prefix + '.gn1.bias'], groups))
    if updown:
        rs = avgpool2 if updown == 'down' else upsample2
        h, x = rs(h), rs(x)
    h = conv3x3(h, P[prefix + '.cv3_1.weight'], P[prefix + '.cv3_1.bias'])
    e = F.linear(silu(emb), P[prefix + '.l_emb.weight'], P[prefix + '.l_emb.bias'])
    if e.shape[1] == 2 * h.shape[1]:   # use_scale_shift_norm (dev/resblock.py:243-247): gn(h) * (1 + scale) + shift
        scale, shift = torch.chunk(e[:, :, None, None], 2, dim=1)
        h = groupnorm(h, P[prefix + '.gn2.weight'], P[prefix + '.gn2.bias'], groups) * (1 + scale) + shift
    else:
        h = groupnorm(h + e[:, :, None, None], P[prefix + '.gn2.weight'], P[prefix + '.gn2.bias'], groups)
    h = silu(h)
    if drop_mask is not None:   # nn.Dropout(p) in training mode with a GIVEN keep-mask: kept values are scaled by 1 / (1 - p)
        h = h * drop_mask.to(h.dtype) / (1.0 - drop_p)
    h = conv3x3(h, P[prefix + '.cv3_2.weight'], P[prefix + '.cv3_2.bias'])
    if (prefix + '.skip_connection.weight') in P:
        x = conv1x1(x, P[prefix + '.skip_connection.weight'], P[prefix + '.skip_connection.bias'])
    return x + h


def mha(qkv: torch.Tensor, n_heads: int) -> torch.Tensor:
    """QKVAttention (dev/unet.py:66-88): qkv (B, 3*C, T) with channel order [Q | K | V], each [NH][HS];
    softmax(q^T k / sqrt(HS)) v.  Same maths as attention_forward1 (train_unet.cu:2553-2616)."""
    B, width, T = qkv.shape
    C = width // 3
    hs = C // n_heads
    q, k, v = qkv.chunk(3, dim=1)
    q = q.reshape(B * n_heads, hs, T)
    k = k.reshape(B * n_heads, hs, T)
    v = v.reshape(B * n_heads, hs, T)
    att = torch.softmax(torch.einsum('bct,bcs->bts', q, k) / math.sqrt(hs), dim=-1)
    return torch.einsum('bts,bcs->bct', att, v).reshape(B, C, T)


def attention_block(x, P, prefix: str, head_size=32, groups=32):
    """dev/unet.py:44-58, train_unet.cu:2933-2953: GN -> qkv 1x1 -> MHA -> proj 1x1 -> + x."""
    B, C, H, W = x.shape
    h = groupnorm(x, P[prefix + '.gn.weight'], P[prefix + '.gn.bias'], groups).reshape(B, C, H * W)
    qkv = F.conv1d(h, P[prefix + '.qkv.weight'], P[prefix + '.qkv.bias'])
    a = mha(qkv, C // head_size)
    return x + F.conv1d(a, P[prefix + '.proj.weight'], P[prefix + '.proj.bias']).reshape(B, C, H, W)


def unet_forward(cfg: UNetConfig, P: Dict[str, torch.Tensor], x: torch.Tensor, t: torch.Tensor,
                 y: torch.Tensor | None = None, drop_masks=None) -> torch.Tensor:
    """dev/unet.py:283-319, train_unet.cu:4335-4416.  `y` (B,) int64 class labels iff cfg.num_classes (dev/unet.py:291-303)."""
    assert (y is not None) == bool(cfg.num_classes), "must specify y if and only if the model is class-conditional"
    emb = timestep_embedding(t, cfg.model_channels, cfg.max_period).to(x.dtype)
    emb = F.linear(emb, P['time_embed.0.weight'], P['time_embed.0.bias'])
    emb = F.linear(silu(emb), P['time_embed.2.weight'], P['time_embed.2.bias'])
    if cfg.num_classes:
        emb = emb + P['label_emb.weight'][y.long()]
    hs: List[torch.Tensor] = []
    h = x
    # drop_masks: one keep-mask (B, C, H, W) per ResBlock in forward order (training with cfg.dropout; None = inference)
    dm = iter(drop_masks) if drop_masks is not None else None
    for l in build_layers(cfg):
        if l.kind in ('linear', 'embed'):
            continue
        if l.kind == 'conv3':
            h = conv3x3(h, P[l.name + '.weight'], P[l.name + '.bias'])
        elif l.kind == 'res':
            h = resblock(h, emb, P, l.name, cfg.gn_groups, drop_mask=next(dm) if dm else None, drop_p=cfg.dropout)
        elif l.kind in ('res_down', 'res_up'):
            h = resblock(h, emb, P, l.name, cfg.gn_groups, updown=l.kind[4:], drop_mask=next(dm) if dm else None,
                         drop_p=cfg.dropout)
        elif l.kind == 'attn':
            h = attention_block(h, P, l.name, cfg.head_size, cfg.gn_groups)
        elif l.kind == 'down':
            h = avgpool2(h)
        elif l.kind == 'up':
            h = upsample2(h)
        elif l.kind == 'concat':
            h = torch.cat([h, hs.pop()], dim=1)
        elif l.kind == 'gn':
            h = silu(groupnorm(h, P[l.name + '.weight'], P[l.name + '.bias'], cfg.gn_groups))
        if l.skip_push:
            hs.append(h)
    assert not hs
    return h


# ----------------------------------------------------------------------------------------------- diffusion
def linear_betas(T: int = 1000) -> np.ndarray:
    """train_unet.py:811-826, train_unet.cu:3131-3147: linear 1e-4 .. 0.02 (float64 linspace cast to float32)."""
    scale = 1000 / T
    return np.linspace(scale * 0.0001, scale * 0.02, T, dtype=np.float64).astype(np.float32)


def diffusion_tables(T: int = 1000):
    """train_unet.py:875-892: float32 cumprod of (1 - beta)."""
    betas = linear_betas(T)
    ac = np.cumprod(1.0 - betas, axis=0)
    return np.sqrt(ac).astype(np.float32), np.sqrt(1.0 - ac).astype(np.float32)


def q_sample(x0: torch.Tensor, t: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """train_unet.py:894-912, train_unet.cu:3200-3215: x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) eps."""
    sa, sb = diffusion_tables()
    idx = t.reshape(-1).long()
    a = torch.from_numpy(sa)[idx].to(x0.dtype).reshape(-1, 1, 1, 1)
    b = torch.from_numpy(sb)[idx].to(x0.dtype).reshape(-1, 1, 1, 1)
    return a * x0 + b * noise


def ddpm_sample(cfg: UNetConfig, P: Dict[str, torch.Tensor], x: torch.Tensor, t_start: int, t_end: int, noises,
                T: int = 1000) -> torch.Tensor:
    """generate.py:29-52 (sample_next_step) looped as generate.py:75-79 does: for t = t_start .. t_end (descending,
    2 <= t < T) with beta_t = betas[t-1], alpha_t = acp[t-1], alpha_t_1 = acp[t-2];  noises[i] is the N(0,1) draw of
    iteration i (generate.py draws it with torch.randn_like)."""
    betas = torch.tensor(linear_betas(T), dtype=torch.float32)
    acp = torch.tensor(np.cumprod(1.0 - linear_betas(T)), dtype=torch.float32)   # GaussianDiffusion.alphas_cumprod
    with torch.no_grad():
        for i, t in enumerate(range(t_start, t_end - 1, -1)):
            assert 2 <= t < T
            beta_t, alpha_t, alpha_t_1 = betas[t - 1], acp[t - 1], acp[t - 2]
            tt = torch.full((x.shape[0], 1), float(t))
            eps = unet_forward(cfg, P, x, tt)
            mu = (x - (beta_t / torch.sqrt(1 - alpha_t)) * eps) / torch.sqrt(1 - beta_t)
            sigma = torch.sqrt((1 - alpha_t_1) / (1 - alpha_t) * beta_t)
            x = mu + sigma * noises[i]
    return x


def random_flip(x0: torch.Tensor, flips) -> torch.Tensor:
    """train_unet.py:531-532 (ImageDataset.__getitem__): `arr = arr[:, ::-1]` on the HWC image = mirror the width axis
    of every image whose coin came up; `flips` holds the coins (the reference draws them with np.random.rand() < 0.5)."""
    out = x0.clone()
    for b, f in enumerate(flips):
        if int(f):
            out[b] = torch.flip(x0[b], dims=[-1])
    return out


def mse_loss(out: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """train_unet.cu:2981-3030: mean over all elements."""
    return ((out - target) ** 2).mean()


def adamw_step(p, g, m, v, step: int, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8, wd=0.0):
    """train_unet.cu:4720-4736 (adamw_kernel2): lerp moments, bias correction, decoupled weight decay."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mh = m / (1 - b1 ** step)
    vh = v / (1 - b2 ** step)
    p = p - lr * (mh / (vh.sqrt() + eps) + wd * p)
    return p, m, v


def synthetic_batch(cfg: UNetConfig, B: int, seed: int = 1234):
    """SURVEY.md section 8(d) config 1 inputs: x0 ~ U[-1,1], t ~ randint(0,1000), eps ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    x0 = torch.rand(B, cfg.in_channels, cfg.H, cfg.W, generator=g) * 2 - 1
    t = torch.randint(0, 1000, (B, 1), generator=g).float()
    noise = torch.randn(B, cfg.in_channels, cfg.H, cfg.W, generator=g)
    return x0, t, noise


def ema_update(ema: torch.Tensor, p: torch.Tensor, rate: float) -> torch.Tensor:
    """Exponential moving average of the parameters after an optimizer step: guided-diffusion's `update_ema`
    (targ.mul_(rate).add_(src, alpha=1 - rate)), the algorithm behind the `ema_rate` option the reference carries
    (train_unet.py:708) but never applies."""
    return ema * rate + p * (1 - rate)


def perturb_zero_params(cfg: UNetConfig, flat: torch.Tensor, seed: int = 5, std: float = 0.02) -> torch.Tensor:
    """Test helper: N(0, std^2) noise on every tensor that is all-zero at initialisation (the zero_module conv2 / proj
    weights, dev/unet.py:21-27, and all biases PyTorch zero-fills), drawn tensor by tensor in parameter order.  With
    conv2 at zero nothing upstream of it -- the whole embedding path in particular -- receives a gradient, so tests of
    that path start from these weights."""
    g = torch.Generator().manual_seed(seed)
    flat = flat.clone()
    off = 0
    for _, shape in param_spec(cfg):
        n = int(np.prod(shape))
        if not flat[off:off + n].any():
            flat[off:off + n] += std * torch.randn(shape, generator=g).reshape(-1)
        off += n
    return flat


def train_step_grads(cfg: UNetConfig, flat: torch.Tensor, x0, t, noise, y=None, drop_masks=None):
    """One forward + backward of the training step (train_unet.cu:5019-5036): returns loss, out, grads(flat)."""
    flat = flat.detach().clone().requires_grad_(True)
    P = unflatten_params(cfg, flat)
    x_t = q_sample(x0.to(flat.dtype), t, noise.to(flat.dtype))
    out = unet_forward(cfg, P, x_t, t.to(flat.dtype), y, drop_masks)
    loss = mse_loss(out, noise.to(flat.dtype))
    loss.backward()
    return loss.detach(), out.detach(), flat.grad.detach()


def train_steps(cfg: UNetConfig, flat: torch.Tensor, batches, lr=1e-4, wd=0.0, threads: int | None = None):
    """n AdamW steps (train_unet.cu:5019-5037); `batches` = list of (x0, t, noise).  Returns losses, final flat."""
    if threads:
        torch.set_num_threads(threads)
    m = torch.zeros_like(flat)
    v = torch.zeros_like(flat)
    losses = []
    for i, (x0, t, noise) in enumerate(batches, 1):
        loss, _, g = train_step_grads(cfg, flat, x0, t, noise)
        flat, m, v = adamw_step(flat, g, m, v, i, lr=lr, wd=wd)
        losses.append(float(loss))
    return losses, flat
