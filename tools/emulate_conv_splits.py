"""CPU emulation (no GPU) of operand-split schemes for the 5x5 conv on real activations: how much accuracy does each
choice of correction-term format buy?  Products are formed exactly in fp64 from the QUANTISED operands, so only the
operand formats differ between rows (TMEM accumulation effects are not modelled).  Decision aid for the next step after
f16f8: would block-scaled FP4 correction terms (tcgen05 kind::mxf4, 2x the fp8 rate) still hold parity?

  main term   x_hi * w_hi            fp16 x fp16                                   (all schemes)
  f16         no correction
  f16x3       + x_lo * w_hi + x_hi * w_lo   (fp16 operands)
  f16f8       + e4m3(x_lo 2^8) e5m2(w 2^-8) + e4m3(x_hi 2^-4) e5m2(w_lo 2^4)      (what the engine runs)
  f16f4       + mxfp4(x_lo) mxfp4(w) + mxfp4(x_hi) mxfp4(w_lo)                    (e2m1, one power-of-two scale per 32 K-elements)
  f16nv4      same with NVFP4 operands (e2m1, one e4m3 scale per 16 K-elements)
  f16f4/8     mxfp4 for the x_lo term only, the engine's fp8 operands for the w_lo term
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dmpfold_oracle as O  # noqa: E402
import bench  # noqa: E402

torch.manual_seed(0)
sd, _ = bench.load_weights()
orc = O.Oracle(sd)
msa = O.encode_aln(O.read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
taps = {}
orc.fold(msa, iterations=0, minsteps=0, taps=taps)


def q_f16(t):
    return t.to(torch.float16).to(torch.float64)


def q_f8(t, dtype, scale):
    return (t * scale).to(torch.float32).to(dtype).to(torch.float64) / scale


E2M1 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], dtype=torch.float64)


def q_mxfp4(t, dim):
    """e2m1 with one shared power-of-two scale per 32 consecutive elements along `dim` (OCP MX rule: scale = 2^(floor(log2 amax) - 2))."""
    t = t.movedim(dim, -1)
    shp = t.shape
    b = t.reshape(-1, 32)
    amax = b.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    scale = torch.exp2(torch.floor(torch.log2(amax)) - 2.0)
    v = (b / scale).clamp(-6.0, 6.0)
    idx = (v.abs().unsqueeze(-1) - E2M1).abs().argmin(dim=-1)
    q = E2M1[idx] * v.sign() * scale
    return q.reshape(shp).movedim(-1, dim)


def q_nvfp4(t, dim):
    """e2m1 with one e4m3 scale per 16 consecutive elements along `dim` (NVFP4; a per-tensor fp32 scale keeps the e4m3
    scales in range, modelled here by normalising with the tensor's amax first)."""
    g = t.abs().max().clamp_min(1e-300)
    t = (t / g).movedim(dim, -1)
    shp = t.shape
    b = t.reshape(-1, 16)
    sc = (b.abs().amax(dim=1, keepdim=True) / 6.0 * 448.0).to(torch.float32).to(torch.float8_e4m3fn).to(torch.float64) / 448.0
    sc = sc.clamp_min(1e-300)
    v = (b / sc).clamp(-6.0, 6.0)
    idx = (v.abs().unsqueeze(-1) - E2M1).abs().argmin(dim=-1)
    q = E2M1[idx] * v.sign() * sc
    return (q.reshape(shp).movedim(-1, dim)) * g


def conv(xq, wq):
    return F.conv2d(xq, wq, None, padding=2)


for blk, key in ((1, 'stem'), (9, 'block8'), (16, 'block15')):
    x = taps[key].double()                                        # (1, 128, L, L) input of block blk
    w = sd[f'resnet.{blk}.layer1.lin.weight'].double()            # (512, 128, 5, 5)
    ref = conv(x, w)
    scale = ref.abs().max()
    xh, wh = q_f16(x), q_f16(w)
    xl, wl = x - xh, w - wh
    main = conv(xh, wh)
    rows = {'f16': main,
            'f16x3': main + conv(q_f16(xl), wh) + conv(xh, q_f16(wl)),
            'f16f8': main + conv(q_f8(xl, torch.float8_e4m3fn, 256.0), q_f8(w, torch.float8_e5m2, 1 / 256.0))
                          + conv(q_f8(xh, torch.float8_e4m3fn, 1 / 16.0), q_f8(wl, torch.float8_e5m2, 16.0)),
            # K runs over (tap, channel) with the 128 channels contiguous -> blocks of 32 channels (dim 1) for both operands
            'f16f4': main + conv(q_mxfp4(xl, 1), q_mxfp4(w, 1)) + conv(q_mxfp4(xh, 1), q_mxfp4(wl, 1)),
            'f16nv4': main + conv(q_nvfp4(xl, 1), q_nvfp4(w, 1)) + conv(q_nvfp4(xh, 1), q_nvfp4(wl, 1)),
            'f16f4/8': main + conv(q_mxfp4(xl, 1), q_mxfp4(w, 1)) + conv(q_f8(xh, torch.float8_e4m3fn, 1 / 16.0), q_f8(wl, torch.float8_e5m2, 16.0)),
            }
    print(f'block {blk}: |out|max {float(scale):.1f}', flush=True)
    for name, y in rows.items():
        err = y - ref
        print(f'   {name:8s} max rel err {float(err.abs().max() / scale):.2e}   rms rel {float(err.pow(2).mean().sqrt() / scale):.2e}', flush=True)
