"""Per-stage precision triangulation (GPU box): for every stage of one pass, the distance of the fp32 oracle and of
the engine from the fp64 evaluation of the same stage, all teacher-forced with the fp64 inputs rounded to fp32.

    python tools/diag_stages64.py [L N seed]      -> prints a table, writes gpurun_out/diag/stages64_*.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dmpfold_oracle as O  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 300
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
SEED = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sd = O.load_state_dict(os.path.join(ROOT, 'dmpfold2_b200', 'trained_model'))
base = O.encode_aln(O.read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
msa = O.synth_msa_structured(base, L, N, SEED)
msa_t = torch.from_numpy(msa)
o32, o64 = O.Oracle(sd), O.Oracle(sd, dtype=torch.float64)
eng = Engine(sd, 0)
rows = []


def err(name, ref64, a32, eng_out):
    ref64 = ref64.double()
    scale = float(ref64.abs().max())
    rms = float(ref64.pow(2).mean().sqrt())

    def d(x):
        x = x.double().cpu()
        return float((x - ref64).abs().max()), float((x - ref64).pow(2).mean().sqrt())
    m32, r32 = d(a32)
    row = {'stage': name, 'max_abs_ref64': scale, 'rms_ref64': rms, 'ref32_max': m32, 'ref32_rms': r32}
    line = '%-22s |ref|max %.3g rms %.3g   ref32: max %.2e rms %.2e' % (name, scale, rms, m32, r32)
    for k, v in eng_out.items():
        me, re_ = d(v)
        row['eng_%s_max' % k], row['eng_%s_rms' % k] = me, re_
        line += '   eng[%s]: max %.2e rms %.2e (x%.1f of ref32)' % (k, me, re_, re_ / max(r32, 1e-30))
    rows.append(row)
    print(line, flush=True)


with torch.no_grad():
    t0 = time.time()
    # ---- MSA features
    f64 = O.msa_features(msa_t, torch.float64)
    f32 = O.msa_features(msa_t)
    fe = eng.dca(msa).cpu()
    err('dca couplings', f64[..., :441], f32[..., :441], {'': fe[..., :441]})
    err('dca apc channel', f64[..., 441], f32[..., 441], {'': fe[..., 441]})
    # ---- vgru / hgru
    v64 = o64.vgru_last(msa_t)
    v32 = o32.vgru_last(msa_t)
    err('vgru', v64, v32, {'': eng.vgru(msa).cpu()})
    h64 = o64.hgru_out(v64)
    h32 = o32.hgru_out(v64.float())
    err('hgru (tf)', h64, h32, {'': eng.hgru(v64.float()).cpu()})
    print('1-D + features %.1fs' % (time.time() - t0), flush=True)
    # ---- one ResNet pass, teacher-forced with the fp64 inputs
    m1_64 = h64.t().contiguous()                                  # (512,L)
    dmap = torch.full((1, 1, L, L), -1.0, dtype=torch.float64)
    x2_64 = torch.cat((f64.permute(2, 0, 1).unsqueeze(0), dmap), dim=1)
    resinp64 = torch.cat(((m1_64.unsqueeze(1) * m1_64.unsqueeze(2)).unsqueeze(0), x2_64), dim=1)
    t0 = time.time()
    taps64 = {}
    sd64 = o64.sd
    head64 = O.resnet_pass(resinp64, sd64, taps64)
    print('fp64 resnet pass %.1fs' % (time.time() - t0), flush=True)
    m1_32 = m1_64.float()
    x2_32 = x2_64.float()
    resinp32 = torch.cat(((m1_32.unsqueeze(1) * m1_32.unsqueeze(2)).unsqueeze(0), x2_32), dim=1)
    taps32 = {}
    head32 = O.resnet_pass(resinp32, o32.sd, taps32)
    heads = {}
    for mode in ('ffma', 'f16x3', 'f16f8'):
        eng.set_conv_mode(mode)
        heads[mode] = eng.resnet_pass(h64.float(), f64.float(), dmap[0, 0].float()).cpu()
    err('resnet pass: head dm', head64[0, 0], head32[0, 0], {k: v[0] for k, v in heads.items()})
    err('resnet pass: head conf', head64[0, 1], head32[0, 1], {k: v[1] for k, v in heads.items()})
    # ---- single blocks, teacher-forced with the fp64 block input
    for k in (1, 8, 16):
        xin = taps64['stem'] if k == 1 else taps64['block%d' % (k - 1)]
        y64 = taps64['block%d' % k]
        y32 = O.resnet_block(xin.float(), o32.sd, k)
        outs = {}
        for mode in ('ffma', 'f16x3', 'f16f8'):
            eng.set_conv_mode(mode)
            outs[mode] = eng.resblock(k, xin[0].permute(1, 2, 0).float().contiguous()).cpu().permute(2, 0, 1)
        err('block %d (tf)' % k, y64[0], y32[0], outs)
    # ---- head -> M -> top-8 MDS, teacher-forced with the fp64 head
    conf64, m64, mds64 = O.head_to_mds(head64)
    conf32, m32, mds32 = O.head_to_mds(head64.float())
    ce, me, mdse = eng.head_mds(head64[0].float())
    err('M (tf)', m64[0], m32[0], {'': me.cpu()})
    err('mds top-8 (tf)', mds64[0], mds32[0], {'': mdse.cpu()})
    # ---- coordinate GRU, teacher-forced
    ca64 = o64.coord_head(m1_64, mds64[0])
    ca32 = o32.coord_head(m1_32, mds64[0].float())
    err('coord gru (tf)', ca64, ca32, {'': eng.coord_gru(h64.float(), mds64[0].float()).cpu()})
    # ---- minimiser, teacher-forced
    r64 = O.refine_coords(ca64, 100)
    r32 = O.refine_coords(ca64.float(), 100)
    err('refine 100 (tf)', r64, r32, {'': eng.refine(ca64.float(), 100).cpu()})

os.makedirs(os.path.join(ROOT, 'gpurun_out', 'diag'), exist_ok=True)
with open(os.path.join(ROOT, 'gpurun_out', 'diag', 'stages64_L%d_N%d_s%d.json' % (L, N, SEED)), 'w') as fh:
    json.dump(rows, fh, indent=1)
eng.close()
