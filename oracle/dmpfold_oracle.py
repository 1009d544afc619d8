"""CPU oracle for the DMPfold2 inference hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a from-scratch CPU restatement (torch CPU ops + numpy) of the algorithm that
`/root/reference/dmpfold/predict.py` and `/root/reference/dmpfold/network.py` execute between the `.aln`
bytes and the (L,5,3) backbone tensor.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product (`dmpfold2_b200/`) never does and fails
loudly when its CUDA library is missing.

Where the arithmetic lives: the reference has no native code; every op dispatches into PyTorch (pinned
torch==1.8.0 upstream, torch 2.11 here: oneDNN conv, MKL GEMM/LAPACK).  The restatement therefore uses the
same torch CPU primitives (conv2d, matmul, linalg) but none of the reference's modules or code.

Parity pinning: the reference's own tests hold NO golden vectors for this path (CI only checks exit codes,
`.github/workflows/CI.yml:31-36`).  The oracle is pinned instead against outputs of the reference itself,
run in the build container by `oracle/make_golden.py` (which imports `/root/reference` with the
`torch.symeig` shim below) and committed under `tests/golden/`; `tests/test_oracle.py` checks the oracle
against those fixtures.  "Oracle B" = reference + canonical eigenvector sign (largest-|component| positive),
because LAPACK's arbitrary signs make the unmodified reference irreproducible across thread counts
(SURVEY.md section 0, fact 5).

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------------
# host prep (predict.py:100-134)
# ----------------------------------------------------------------------------------------------------
_AA_FROM = 'ARNDCQEGHILKMFPSTWYVBJOUXZ-.'
_AA_TO = 'ABCDEFGHIJKLMNOPQRSTUUUUUUVV'
MAX_SEQS = 3000


def read_aln(path: str) -> List[str]:
    """predict.py:100-104 -- every line not starting with '>' is a row, rstrip()'d."""
    rows = []
    with open(path, 'r') as fh:
        for line in fh.readlines():
            if not line.startswith('>'):
                rows.append(line.rstrip())
    return rows


def encode_aln(rows: List[str]) -> np.ndarray:
    """predict.py:124-132 -- residue letters -> codes 0..19, BJOUXZ -> 20, '-.' -> 21; N capped at 3000."""
    trans = str.maketrans(_AA_FROM, _AA_TO)
    nseqs, length = len(rows), len(rows[0])
    mat = (np.frombuffer(''.join(rows).translate(trans).encode('latin-1'), dtype=np.uint8) - ord('A'))
    mat = mat.reshape(nseqs, length)
    if nseqs > MAX_SEQS:
        mat = mat[:MAX_SEQS]
    return mat


def read_template_ca(path: str) -> np.ndarray:
    """predict.py:106-117 -- all ATOM records whose name field is ' CA ', fixed columns 30:54."""
    out = []
    with open(path, 'r') as fh:
        for line in fh:
            if line[:4] == 'ATOM' and line[12:16] == ' CA ':
                out.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return np.asarray(out, dtype=np.float32)


# ----------------------------------------------------------------------------------------------------
# MSA features (predict.py:32-61)
# ----------------------------------------------------------------------------------------------------
def one_hot_msa(msa: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """predict.py:136 -- one_hot(clamp(codes, max=20), 21): gap (21) and unknown (20) share class 20."""
    return F.one_hot(torch.clamp(msa.long(), max=20), 21).to(dtype)


def reweight(msa1hot: torch.Tensor, cutoff: float = 0.8) -> torch.Tensor:
    """predict.py:32-37 -- w_n = 1 / #{m : identical columns(n,m) > cutoff*L}; strict '>' evaluated in f32."""
    id_min = msa1hot.shape[1] * cutoff
    n, l, a = msa1hot.shape
    x = msa1hot.reshape(n, l * a)
    id_mtx = x @ x.t()
    return 1.0 / (id_mtx > id_min).to(msa1hot.dtype).sum(dim=-1)


def fast_dca(msa1hot: torch.Tensor, weights: torch.Tensor, penalty: float = 4.5) -> torch.Tensor:
    """predict.py:41-61 -- shrunk weighted covariance, inverse, (L,L,441) couplings + APC contact channel."""
    nr, nc, ns = msa1hot.shape
    x = msa1hot.reshape(nr, nc * ns)
    num_points = weights.sum() - torch.sqrt(weights.mean())
    mean = (x * weights[:, None]).sum(dim=0, keepdim=True) / num_points
    x = (x - mean) * torch.sqrt(weights[:, None])
    cov = (x.t() @ x) / num_points
    cov_reg = cov + torch.eye(nc * ns, device=x.device, dtype=x.dtype) * penalty / torch.sqrt(weights.sum())
    inv_cov = torch.linalg.inv(cov_reg)
    x1 = inv_cov.view(nc, ns, nc, ns)
    features = x1.transpose(1, 2).contiguous().reshape(nc, nc, ns * ns)
    eye = torch.eye(nc, device=x.device, dtype=x.dtype)
    x3 = torch.sqrt((x1[:, :-1, :, :-1] ** 2).sum(dim=(1, 3))) * (1 - eye)
    apc = x3.sum(dim=0, keepdim=True) * x3.sum(dim=1, keepdim=True) / x3.sum()
    contacts = (x3 - apc) * (1 - eye)
    return torch.cat((features, contacts[:, :, None]), dim=2)


def msa_features(msa: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """predict.py:136-140 -- (L,L,442) DCA features, zeros for a single sequence.
    dtype=float64 is the ground-truth arm of the precision triangulation (tools/fp64_triangulate.py)."""
    nseqs, length = msa.shape
    hot = one_hot_msa(msa, dtype)
    w = reweight(hot, 0.8)
    if nseqs > 1:
        return fast_dca(hot, w).to(dtype)
    return torch.zeros((length, length, 442), device=msa.device, dtype=dtype)


# ----------------------------------------------------------------------------------------------------
# GRU (torch.nn.GRU semantics; gate order r,z,n) -- network.py:189-190, :213
# ----------------------------------------------------------------------------------------------------
def gru_cell(gi: torch.Tensor, gh: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """One GRU step from the two pre-activations: r=s(gi_r+gh_r), z=s(gi_z+gh_z), n=tanh(gi_n+r*gh_n)."""
    hs = h.shape[-1]
    r = torch.sigmoid(gi[..., :hs] + gh[..., :hs])
    z = torch.sigmoid(gi[..., hs:2 * hs] + gh[..., hs:2 * hs])
    n = torch.tanh(gi[..., 2 * hs:] + r * gh[..., 2 * hs:])
    return (1 - z) * n + z * h


def gru_layer(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh, reverse: bool = False) -> torch.Tensor:
    """x (T,B,I) -> (T,B,H); h0 = 0.  Explicit time loop (the small-case restatement of nn.GRU)."""
    t_len, b, _ = x.shape
    hs = w_hh.shape[1]
    gi_all = x @ w_ih.t() + b_ih
    h = torch.zeros((b, hs), device=x.device)
    out = [None] * t_len
    order = range(t_len - 1, -1, -1) if reverse else range(t_len)
    for t in order:
        gh = h @ w_hh.t() + b_hh
        h = gru_cell(gi_all[t], gh, h)
        out[t] = h
    return torch.stack(out, dim=0)


def gru_stack(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, layers: int, bidir: bool) -> torch.Tensor:
    """Multi-layer (bi)GRU over x (T,B,I) with state_dict naming `prefix.weight_ih_l{k}[_reverse]`."""
    for k in range(layers):
        outs = []
        for suffix, rev in (('', False), ('_reverse', True)) if bidir else (('', False),):
            outs.append(gru_layer(x, sd[f'{prefix}.weight_ih_l{k}{suffix}'], sd[f'{prefix}.weight_hh_l{k}{suffix}'],
                                  sd[f'{prefix}.bias_ih_l{k}{suffix}'], sd[f'{prefix}.bias_hh_l{k}{suffix}'], rev))
        x = torch.cat(outs, dim=-1)
    return x


def _nn_gru(sd: Dict[str, torch.Tensor], prefix: str, inp: int, hid: int, layers: int, bidir: bool) -> torch.nn.GRU:
    """torch.nn.GRU loaded with the reference weights -- the same ATen kernel the reference runs
    (used for the full-size oracle and the CPU baseline timing; gru_stack above is its restatement)."""
    g = torch.nn.GRU(inp, hid, num_layers=layers, bidirectional=bidir)
    g.load_state_dict({k[len(prefix) + 1:]: v.cpu() for k, v in sd.items() if k.startswith(prefix + '.')})
    return g.eval()


# ----------------------------------------------------------------------------------------------------
# ResNet pieces (network.py:12-103)
# ----------------------------------------------------------------------------------------------------
def maxout_norm(x: torch.Tensor, w, b, gamma, beta, pool: int, pad: int) -> torch.Tensor:
    """network.py:25-34 -- conv -> max over `pool` consecutive channels -> InstanceNorm(affine, eps 1e-5)."""
    y = F.conv2d(x, w, b, padding=pad)
    n, c, h, wd = y.shape
    y = y.view(n, c // pool, pool, h, wd).max(dim=2)[0]
    return F.instance_norm(y, weight=gamma, bias=beta, eps=1e-5)


def scse(y: torch.Tensor, fc0, fc2, sse_w, sse_b) -> torch.Tensor:
    """network.py:37-81 -- y*sigmoid(W2 relu(W1 avgpool(y))) + y*sigmoid(conv1x1(y))."""
    n, c, _, _ = y.shape
    g = y.mean(dim=(2, 3))
    g = torch.sigmoid(F.linear(F.relu(F.linear(g, fc0)), fc2)).view(n, c, 1, 1)
    s = torch.sigmoid(F.conv2d(y, sse_w, sse_b))
    return y * g + y * s


def resnet_block(x: torch.Tensor, sd: Dict[str, torch.Tensor], k: int) -> torch.Tensor:
    """network.py:94-103 (dropouts are identity in eval)."""
    p = f'resnet.{k}'
    y = maxout_norm(x, sd[f'{p}.layer1.lin.weight'], sd[f'{p}.layer1.lin.bias'], sd[f'{p}.layer1.norm.weight'],
                    sd[f'{p}.layer1.norm.bias'], 4, 2)
    y = scse(y, sd[f'{p}.scSE.cSE.fc.0.weight'], sd[f'{p}.scSE.cSE.fc.2.weight'],
             sd[f'{p}.scSE.sSE.conv.weight'], sd[f'{p}.scSE.sSE.conv.bias'])
    return y + x


def resnet_pass(resinp: torch.Tensor, sd: Dict[str, torch.Tensor], taps: Optional[dict] = None) -> torch.Tensor:
    """network.py:194-209,235 -- stem (1x1, maxout 3) -> 16 blocks -> 1x1 conv to 2 channels."""
    x = maxout_norm(resinp, sd['resnet.0.lin.weight'], sd['resnet.0.lin.bias'], sd['resnet.0.norm.weight'],
                    sd['resnet.0.norm.bias'], 3, 0)
    if taps is not None:
        taps['stem'] = x
    for k in range(1, 17):
        x = resnet_block(x, sd, k)
        if taps is not None:
            taps[f'block{k}'] = x
    return F.conv2d(x, sd['resnet.17.weight'], sd['resnet.17.bias'])


# ----------------------------------------------------------------------------------------------------
# head, MDS, coordinate GRU (network.py:237-255)
# ----------------------------------------------------------------------------------------------------
def canonical_sign(v: torch.Tensor) -> torch.Tensor:
    """Oracle-B rule: for every eigenvector (column) make its largest-|component| entry positive
    (lowest index wins ties, as torch.argmax does)."""
    idx = v.abs().argmax(dim=-2, keepdim=True)
    return v * torch.gather(v, -2, idx).sign()


def symeig_canonical(m: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Replacement for the removed torch.symeig(M, eigenvectors=True) (network.py:247): ascending
    eigenvalues from the upper triangle, with the canonical sign rule."""
    w, v = torch.linalg.eigh(m, UPLO='U')
    return w, canonical_sign(v)


def head_to_mds(x2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """network.py:237-250 -- x2 (1,2,L,L) -> conf (1,L), M (1,L,L), mds (1,L,8)."""
    nres = x2.shape[-1]
    dm = x2[:, 0]
    conf = x2[:, 1].mean(dim=2)
    dm = torch.abs((dm + dm.transpose(1, 2)) / 2)
    m = 0.5 * (dm[:, 0:1, :].expand(-1, nres, -1) ** 2 + dm[:, :, 0:1].expand(-1, -1, nres) ** 2 - dm ** 2)
    w, v = symeig_canonical(m if m.dtype == torch.float64 else m.float())
    w = torch.clamp(F.relu(w), min=1e-8)
    mds = torch.matmul(v, torch.diag_embed(w.sqrt()))[:, :, -8:]
    return conf, m, mds


# ----------------------------------------------------------------------------------------------------
# geometry (network.py:106-177)
# ----------------------------------------------------------------------------------------------------
def refine_coords(coords: torch.Tensor, n_steps: int) -> torch.Tensor:
    """network.py:106-137 -- explicit steric (3.0 A) + bond (3.78 A) force descent on the CA trace (L,3)."""
    for _ in range(n_steps):
        diffs = coords.unsqueeze(0) - coords.unsqueeze(1)          # [i,j] = c_j - c_i
        dists = diffs.norm(dim=2).clamp(min=0.01, max=10.0)
        viol = (dists < 3.0).to(coords.dtype) * (3.0 - dists)
        accels = ((100.0 * viol).unsqueeze(2) * (diffs / dists.unsqueeze(2))).sum(dim=0)
        d = coords[1:] - coords[:-1]
        dd = d.norm(dim=1).clamp(min=0.1)
        acov = (100.0 * (dd - 3.78).clamp(max=3.0)).unsqueeze(1) * (d / dd.unsqueeze(1))
        accels = accels.clone()
        accels[:-1] += acov
        accels[1:] -= acov
        coords = coords + accels.clamp(min=-100.0, max=100.0) * 0.001
    return coords


def calpha_to_main_chain(ca: torch.Tensor) -> torch.Tensor:
    """network.py:141-177 -- CA trace (1,L,3) -> N,CA,C,O,CB (1,5L,3) by the Levitt/Taylor construction."""
    def unit(v):
        return F.normalize(v, dim=2)
    n1 = ca[:, :1] - ca[:, 1:2]
    n3 = ca[:, 2:3] - ca[:, 1:2]
    c1 = ca[:, -1:] - ca[:, -2:-1]
    c3 = ca[:, -3:-2] - ca[:, -2:-1]
    ext = torch.cat((ca[:, :1] + 3.82 * unit(torch.cross(n1, n3, dim=2)), ca,
                     ca[:, -1:] + 3.82 * unit(torch.cross(c1, c3, dim=2))), dim=1)
    v_n = ext[:, :-2] - ext[:, 1:-1]
    v_c = ext[:, 2:] - ext[:, 1:-1]
    mid = (ext[:, 1:] + ext[:, :-1]) / 2
    x = unit(torch.cross(v_n, v_c, dim=2))
    at_n = mid[:, :-1] - v_n / 8 + x / 4
    c_shift = mid[:, :-1] + v_n / 8 - x / 2
    o_shift = mid[:, :-1] - x * 1.8
    c_term = mid[:, -1:] - v_c[:, -1:] / 8 + x[:, -1:] / 2
    o_term = mid[:, -1:] + x[:, -1:] * 2.0
    at_c = torch.cat((c_shift[:, 1:], c_term), dim=1)
    at_o = torch.cat((o_shift[:, 1:], o_term), dim=1)
    v_nca = ca - at_n
    v_cca = ca - at_c
    cr = torch.cross(v_nca, v_cca, dim=2)
    bis = v_nca + v_cca
    ang = math.pi / 2 - math.asin(1 / math.sqrt(3))
    sx = (1.5 * math.cos(ang) / bis.norm(dim=2)).unsqueeze(2)
    sy = (1.5 * math.sin(ang) / cr.norm(dim=2)).unsqueeze(2)
    at_cb = ca + sx * bis + sy * cr
    out = torch.stack((at_n, ca, at_c, at_o, at_cb), dim=2)
    return out.reshape(ca.shape[0], 5 * ca.shape[1], 3)


# ----------------------------------------------------------------------------------------------------
# full forward (network.py:218-314) and API (predict.py:74-158)
# ----------------------------------------------------------------------------------------------------
class Oracle:
    """Holds the weights (a state_dict as loaded by predict.py:89-92) and runs the reference algorithm."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: str = 'cpu', dtype: torch.dtype = torch.float32):
        # device='cpu' is the oracle proper.  A cuda device only serves tools/torch_cuda_bar.py, which times the same
        # algorithm on PyTorch's library kernels (cuDNN/cuBLAS/cuSOLVER) as the "existing Blackwell kernels" bar.
        # dtype=float32 is the reference's arithmetic (the parity target); dtype=float64 runs the SAME algorithm with
        # the same (fp32-valued) weights in double precision -- the ground truth both the reference and the engine are
        # measured against in tools/fp64_triangulate.py.
        self.device = torch.device(device)
        self.dtype = dtype
        self.sd = {k: v.detach().float().to(dtype).to(self.device) for k, v in state_dict.items()}
        self._vgru = _nn_gru(self.sd, 'vgru', 22, 512, 2, False).to(self.device).to(dtype)
        self._hgru = _nn_gru(self.sd, 'hgru', 512, 256, 2, True).to(self.device).to(dtype)
        self._cgru = _nn_gru(self.sd, 'coord_gru', 520, 256, 3, True).to(self.device).to(dtype)

    # -- 1-D track: network.py:223-226
    def mat1d(self, msa: torch.Tensor) -> torch.Tensor:
        """(N,L) codes -> (512,L).  embed is the identity (network.py:188) so the input is one_hot(.,22)."""
        x = F.one_hot(msa.long(), 22).to(self.dtype)
        v = self._vgru(x)[0][-1]                                  # (L,512): state after the last MSA row
        h = self._hgru(v.unsqueeze(1))[0]                         # (L,1,512)
        return h.permute(1, 2, 0)[0]                              # (512,L)

    def vgru_last(self, msa: torch.Tensor) -> torch.Tensor:
        return self._vgru(F.one_hot(msa.long(), 22).to(self.dtype))[0][-1]

    def hgru_out(self, v: torch.Tensor) -> torch.Tensor:
        return self._hgru(v.unsqueeze(1))[0][:, 0]

    def coord_head(self, mat1d: torch.Tensor, mds: torch.Tensor) -> torch.Tensor:
        """network.py:251-255 -- (512,L),(L,8) -> CA (L,3)."""
        emb = torch.cat((mat1d.t(), mds), dim=1).unsqueeze(1)     # time-major (L,1,520) == batch_first (1,L,520)
        out = self._cgru(emb)[0][:, 0]
        return out @ self.sd['coord_fc.weight'].t()

    def one_pass(self, resinp: torch.Tensor, mat1d: torch.Tensor, taps: Optional[dict] = None):
        x2 = resnet_pass(resinp, self.sd, taps)
        conf, m, mds = head_to_mds(x2)
        if taps is not None:
            taps['head'] = x2
            taps['M'] = m
            taps['mds'] = mds
        return self.coord_head(mat1d, mds[0]), conf[0]

    def forward(self, msa: torch.Tensor, x2: torch.Tensor, nloops: int, refine_steps: int,
                taps: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """network.py:218-314.  msa (N,L) int, x2 (1,443,L,L).  Returns coords (5L,3), conf (L,)."""
        with torch.no_grad():
            m1 = self.mat1d(msa)
            outer = (m1.unsqueeze(1) * m1.unsqueeze(2)).unsqueeze(0)      # [c,i,j] = m[c,i]*m[c,j]
            resinp = torch.cat((outer, x2), dim=1)
            if taps is not None:
                taps['mat1d'] = m1
            ca, conf = self.one_pass(resinp, m1, taps)
            if taps is not None:
                taps['ca0'] = ca
            if refine_steps > 0:
                ca = refine_coords(ca, refine_steps)
            best_ca, best_conf = ca, conf
            for _ in range(nloops):
                d = ca.unsqueeze(0) - ca.unsqueeze(1)
                dmap = torch.clamp(d.pow(2).sum(dim=2), min=1e-8).sqrt()
                resinp = torch.cat((resinp[:, :-1], dmap[None, None]), dim=1)
                ca, conf = self.one_pass(resinp, m1)
                if conf.mean() > best_conf.mean():
                    best_ca, best_conf = ca, conf
            if refine_steps > 0:
                best_ca = refine_coords(best_ca, refine_steps)
            return calpha_to_main_chain(best_ca.unsqueeze(0))[0], torch.sigmoid(best_conf)

    def fold(self, msa: np.ndarray, template_ca: Optional[np.ndarray] = None, iterations: int = 10,
             minsteps: int = 100, taps: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """predict.py:121-153 given the encoded alignment.  Returns coords (L,5,3), confs (L,)."""
        msa_t = torch.from_numpy(np.ascontiguousarray(msa)).long().to(self.device)
        length = msa_t.shape[1]
        with torch.no_grad():
            feats = msa_features(msa_t, self.dtype).permute(2, 0, 1).unsqueeze(0)
            if template_ca is not None:
                c = torch.from_numpy(template_ca).float().to(self.dtype).unsqueeze(0).to(self.device)
                dmap = (c - c.transpose(0, 1)).pow(2).sum(dim=2).sqrt()[None, None]
            else:
                dmap = torch.zeros((1, 1, length, length), device=self.device, dtype=self.dtype) - 1
            x2 = torch.cat((feats, dmap), dim=1)
            if taps is not None:
                taps['x2'] = x2
        coords, conf = self.forward(msa_t, x2, max(iterations, 0), max(minsteps, 0), taps)
        return coords.view(length, 5, 3), conf


def load_state_dict(weights_dir: str) -> Dict[str, torch.Tensor]:
    """predict.py:83-92 -- the two-part weight file merged by dict.update."""
    import os
    sd = torch.load(os.path.join(weights_dir, 'FINAL_fullmap_e2e_model_part1.pt'), map_location='cpu')
    sd.update(torch.load(os.path.join(weights_dir, 'FINAL_fullmap_e2e_model_part2.pt'), map_location='cpu'))
    return sd


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


def synth_msa_tandem(base: np.ndarray, length: int, nseqs: int, seed: int, mut: float = 0.02) -> np.ndarray:
    """Tandem-repeat synthetic MSA: every column block reuses the SAME resampled rows of `base` (so the covariation
    is consistent across the repeats) with `mut` point mutations; row 0 stays the (tiled) query."""
    rng = np.random.default_rng(seed)
    n0, l0 = base.shape
    rows = np.concatenate(([0], rng.integers(1, n0, size=nseqs - 1)))
    nblk = -(-length // l0)
    msa = np.concatenate([base[rows]] * nblk, axis=1)[:, :length].copy()
    m = rng.random(msa.shape) < mut
    m[0] = False
    msa[m] = rng.integers(0, 20, size=int(m.sum()), dtype=np.uint8)
    return msa


def synth_template_domains(domain_ca: np.ndarray, length: int, gap: float = 8.0) -> np.ndarray:
    """Template CA trace for a tiled synthetic target: copies of one folded domain (e.g. the reference's own PF10963
    prediction, tests/golden/pf10963_n10_m100.npz) laid out along x, `gap` Angstrom apart, cropped to `length` residues.
    Seeding the recycling loop with it (predict.py:142-143) gives a confident, well-conditioned L=300 target."""
    dom = np.asarray(domain_ca, dtype=np.float64)
    span = dom.max(0) - dom.min(0)
    ncopy = -(-length // len(dom))
    tm = np.concatenate([dom + np.array([k * (span[0] + gap), 0.0, 0.0]) for k in range(ncopy)])[:length]
    return tm.astype(np.float32)


def kabsch_rmsd(a: np.ndarray, b: np.ndarray) -> float:
    """RMSD of two (L,3) point sets after optimal superposition."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    a0, b0 = a - a.mean(0), b - b.mean(0)
    u, s, vt = np.linalg.svd(a0.T @ b0)
    d = np.sign(np.linalg.det(u @ vt))
    e = (a0 ** 2).sum() + (b0 ** 2).sum() - 2 * (s[0] + s[1] + d * s[2])
    return float(np.sqrt(max(e, 0.0) / len(a)))
