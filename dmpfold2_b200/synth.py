"""Synthetic inputs for benchmarking and smoke runs (no I/O, numpy only).

`synth_msa_structured` resamples rows/columns of a real alignment (well-conditioned),
`synth_msa_random` draws an i.i.d.-mutation alignment (throughput only), `random_state_dict` builds random weights
of the reference architecture for when the trained files are not available.  These are deliberately independent of
`oracle/` (the test-only CPU restatement), which keeps its own copies for the parity tests.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch


def synth_msa_random(length: int, nseqs: int, seed: int) -> np.ndarray:
    """i.i.d.-mutation synthetic MSA (throughput only -- chaotic input for parity)."""
    rng = np.random.default_rng(seed)
    query = rng.integers(0, 20, size=length, dtype=np.uint8)
    msa = np.tile(query, (nseqs, 1))
    for n in range(1, nseqs):
        p = rng.uniform(0.05, 0.7)
        mut = rng.random(length) < p
        msa[n, mut] = rng.integers(0, 20, size=int(mut.sum()), dtype=np.uint8)
        for _ in range(rng.poisson(0.01 * length)):
            s = rng.integers(0, length)
            msa[n, s:s + rng.integers(1, 10)] = 21
    return msa


def synth_msa_structured(base: np.ndarray, length: int, nseqs: int, seed: int) -> np.ndarray:
    """Well-conditioned synthetic MSA: rows/columns resampled from a real alignment (`base`, e.g. PF10963)
    with 5 % point mutations; row 0 stays the (tiled) query."""
    rng = np.random.default_rng(seed)
    n0, l0 = base.shape
    nblk = -(-length // l0)
    blocks = []
    rows = np.concatenate(([0], rng.integers(1, n0, size=nseqs - 1)))
    for b in range(nblk):
        r = rows.copy()
        if b > 0:
            perm = rng.permutation(np.arange(1, n0))
            r[1:] = perm[(rows[1:] - 1) % (n0 - 1)]
        blocks.append(base[r])
    msa = np.concatenate(blocks, axis=1)[:, :length].copy()
    mut = rng.random(msa.shape) < 0.05
    mut[0] = False
    msa[mut] = rng.integers(0, 20, size=int(mut.sum()), dtype=np.uint8)
    return msa


def random_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random weights of the reference architecture/shapes for when the trained
    files are unavailable; magnitudes chosen so activations stay in a realistic range."""
    g = torch.Generator().manual_seed(seed)

    def u(*shape, scale):
        return (torch.rand(*shape, generator=g) * 2 - 1) * scale
    sd = {'embed.weight': torch.eye(22)}

    def gru(prefix, inp, hid, layers, bidir):
        for k in range(layers):
            for suf in (('', '_reverse') if bidir else ('',)):
                i = inp if k == 0 else hid * (2 if bidir else 1)
                s = 1.0 / math.sqrt(hid)
                sd[f'{prefix}.weight_ih_l{k}{suf}'] = u(3 * hid, i, scale=s)
                sd[f'{prefix}.weight_hh_l{k}{suf}'] = u(3 * hid, hid, scale=s)
                sd[f'{prefix}.bias_ih_l{k}{suf}'] = u(3 * hid, scale=s)
                sd[f'{prefix}.bias_hh_l{k}{suf}'] = u(3 * hid, scale=s)
    gru('vgru', 22, 512, 2, False)
    gru('hgru', 512, 256, 2, True)
    sd['resnet.0.lin.weight'] = u(384, 955, 1, 1, scale=0.07)
    sd['resnet.0.lin.bias'] = u(384, scale=0.03)
    sd['resnet.0.norm.weight'] = 1 + u(128, scale=0.2)
    sd['resnet.0.norm.bias'] = u(128, scale=0.2)
    for k in range(1, 17):
        p = f'resnet.{k}'
        sd[f'{p}.layer1.lin.weight'] = u(512, 128, 5, 5, scale=0.03)
        sd[f'{p}.layer1.lin.bias'] = u(512, scale=0.02)
        sd[f'{p}.layer1.norm.weight'] = 1 + u(128, scale=0.2)
        sd[f'{p}.layer1.norm.bias'] = u(128, scale=0.2)
        sd[f'{p}.scSE.cSE.fc.0.weight'] = u(8, 128, scale=0.2)
        sd[f'{p}.scSE.cSE.fc.2.weight'] = u(128, 8, scale=0.4)
        sd[f'{p}.scSE.sSE.conv.weight'] = u(1, 128, 1, 1, scale=0.1)
        sd[f'{p}.scSE.sSE.conv.bias'] = u(1, scale=0.1)
    sd['resnet.17.weight'] = u(2, 128, 1, 1, scale=0.3)
    sd['resnet.17.bias'] = torch.tensor([8.0, 0.0])
    gru('coord_gru', 520, 256, 3, True)
    sd['coord_fc.weight'] = u(3, 512, scale=1.5)
    return sd


# ----------------------------------------------------------------------------------------------------
# synthetic inputs and metrics (section A.4)
# ----------------------------------------------------------------------------------------------------
