"""Host side of the B200-native DMPfold2 engine: the drop-in for the reference's `dmpfold.predict`.

Same public surface as the reference (dmpfold/predict.py:74-75, :160-208):
    aln_to_coords(input_file, device='cpu', template=None, iterations=10, minsteps=100,
                  weights_file=None, return_alnmat=False) -> (coords (L,5,3), confs (L,)[, alnmat])
    run_dmpfold()   -- the `dmpfold` CLI, flags -i -d -t -n -m -w, PDB text on stdout.

What differs by design: the arithmetic between the encoded alignment and the (L,5,3) tensor runs in
libdmp2.so (hand-written sm_100a kernels) instead of a PyTorch module graph, and the weights are loaded and
repacked once per (weights file, device) instead of on every call (reference: predict.py:79-98).

Device semantics: there is NO CPU execution path.  `device` names where the returned tensors live, as in the
reference; the compute always runs on a CUDA device -- the one named by `device` if it is a cuda device,
else cuda:$DMPFOLD_CUDA_DEVICE (default 0).  With no CUDA device the call raises.
"""
from __future__ import print_function

import argparse
import os
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .engine import Engine

default_device = 'cpu'
default_iterations = 10
default_minsteps = 100

MAX_SEQS = 3000                                     # predict.py:130-132
_AA_TRANS = str.maketrans('ARNDCQEGHILKMFPSTWYVBJOUXZ-.', 'ABCDEFGHIJKLMNOPQRSTUUUUUUVV')   # predict.py:124

_ENGINES: Dict[Tuple[str, int, int], Engine] = {}


def default_weights_dir() -> str:
    return os.environ.get('DMPFOLD_WEIGHTS_DIR', os.path.join(os.path.dirname(os.path.realpath(__file__)), 'trained_model'))


MODEL_URL = 'https://github.com/psipred/DMPfold2/raw/master/dmpfold/trained_model/FINAL_fullmap_e2e_model_part{part}.pt'


def download_trained_model(modeldir: str) -> None:
    """predict.py:64-71 -- first-time download of the two weight parts (~140 MB) into `modeldir`."""
    from urllib import request
    print('Downloading trained model (~140 MB) as first time setup to ', modeldir, ', internet connection required',
          sep='', file=sys.stderr)
    if not os.path.isdir(modeldir):
        os.mkdir(modeldir)
    for part in ['1', '2']:
        request.urlretrieve(MODEL_URL.format(part=part), os.path.join(modeldir, f'FINAL_fullmap_e2e_model_part{part}.pt'))


def load_weights(weights_file: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """predict.py:83-96 -- two-part default weights merged by dict.update, or one custom file (-w); downloaded on
    first use like the reference does when they are not there yet."""
    if weights_file is None:
        d = default_weights_dir()
        p1 = os.path.join(d, 'FINAL_fullmap_e2e_model_part1.pt')
        p2 = os.path.join(d, 'FINAL_fullmap_e2e_model_part2.pt')
        if not os.path.isfile(p1) or not os.path.isfile(p2):
            try:
                download_trained_model(d)
            except Exception as ex:
                raise FileNotFoundError(f'trained model not found in {d} (expected FINAL_fullmap_e2e_model_part[12].pt) and the '
                                        f'download failed ({ex}); copy the two files there, set DMPFOLD_WEIGHTS_DIR or pass '
                                        'weights_file') from ex
        sd = torch.load(p1, map_location='cpu')
        sd.update(torch.load(p2, map_location='cpu'))
        return sd
    return torch.load(weights_file, map_location='cpu')


def get_engine(weights_file: Optional[str], device_index: int) -> Engine:
    """One engine per (weights, device, CUDA stream): an engine owns ONE workspace and its folds are asynchronous on the
    caller's current stream, so callers on different streams (or threads with their own streams) must not share one."""
    stream = torch.cuda.current_stream(torch.device('cuda', device_index)).cuda_stream if torch.cuda.is_available() else 0
    key = (os.path.realpath(weights_file) if weights_file else '<default>', device_index, int(stream))
    eng = _ENGINES.get(key)
    if eng is None:
        eng = Engine(load_weights(weights_file), device_index)
        _ENGINES[key] = eng
    return eng


def read_aln(input_file: str) -> List[str]:
    """predict.py:100-104."""
    aln = []
    with open(input_file, 'r') as alnfile:
        for line in alnfile.readlines():
            if not line.startswith('>'):
                aln.append(line.rstrip())
    return aln


def a3m_to_aln(a3m_path: str, aln_path: str) -> int:
    """The README's `grep -v '^>' file.a3m | sed -e 's/[a-z]//g' > file.aln` (reference README.md:30-33): drop
    header lines and the lower-case insertion columns of an hhblits a3m.  Returns the number of sequences."""
    import re
    n = 0
    with open(a3m_path, 'r') as fi, open(aln_path, 'w') as fo:
        for line in fi:
            if line.startswith('>'):
                continue
            fo.write(re.sub('[a-z]', '', line))
            n += 1
    return n


def encode_aln(aln: List[str]) -> np.ndarray:
    """predict.py:124-132 -- uint8 (N,L) residue codes; N truncated to the first 3000 rows."""
    nseqs, length = len(aln), len(aln[0])
    alnmat = (np.frombuffer(''.join(aln).translate(_AA_TRANS).encode('latin-1'), dtype=np.uint8) - ord('A')).reshape(nseqs, length)
    if nseqs > MAX_SEQS:
        alnmat = alnmat[:MAX_SEQS]
    return alnmat


def read_template(template: str) -> np.ndarray:
    """predict.py:106-117 -- every ATOM record named ' CA ', fixed-column xyz."""
    coords = []
    with open(template, 'r') as fh:
        for line in fh:
            if line[:4] == 'ATOM' and line[12:16] == ' CA ':
                coords.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return np.asarray(coords, dtype=np.float32).reshape(-1, 3)


def _compute_device(device: torch.device) -> int:
    if device.type == 'cuda':
        return device.index if device.index is not None else torch.cuda.current_device()
    return int(os.environ.get('DMPFOLD_CUDA_DEVICE', '0'))


def aln_to_coords(input_file, device=default_device, template=None, iterations=default_iterations,
                  minsteps=default_minsteps, weights_file=None, return_alnmat=False):
    device = torch.device(device)
    engine = get_engine(weights_file, _compute_device(device))

    alnmat = encode_aln(read_aln(input_file))
    length = alnmat.shape[1]
    if np.any(alnmat > 21):
        raise ValueError('alignment contains characters outside the residue alphabet')
    init_coords = None
    if template is not None:
        init_coords = read_template(template)
        if init_coords.shape[0] != length:      # the reference fails inside torch.cat (predict.py:147)
            raise RuntimeError(f'Sizes of tensors must match: template has {init_coords.shape[0]} CA atoms, '
                               f'alignment has {length} columns')
        init_coords = torch.from_numpy(init_coords)

    inputs = torch.from_numpy(np.ascontiguousarray(alnmat))
    coords, confs = engine.fold(inputs, init_coords, max(iterations, 0), max(minsteps, 0))
    coords, confs = coords.to(device), confs.to(device)
    if return_alnmat:
        return coords, confs, alnmat
    return coords, confs


_POOLS: Dict[Tuple[str, int, int], object] = {}


def alns_to_coords(input_files, device='cuda', templates=None, iterations=default_iterations, minsteps=default_minsteps,
                   weights_file=None, gather=True, streams=1):
    """Batch form of aln_to_coords for many independent alignments (BASELINE.json configs[2]).  Under
    torch.distributed (one process per GPU) the list is sharded round-robin over the ranks with no data-path
    collective (dmpfold2_b200.parallel); results come back as CPU tensors in input order.
    streams > 1: throughput mode -- every rank keeps `streams` targets in flight on its GPU (one engine + one CUDA stream
    each, parallel.StreamPool), so the latency-bound stages of one target run beside the convs of another."""
    from .parallel import StreamPool, fold_many, fold_many_batched
    templates = templates or [None] * len(input_files)
    if streams <= 1:
        def one(job):
            path, tmpl = job
            coords, confs = aln_to_coords(path, device=device, template=tmpl, iterations=iterations, minsteps=minsteps,
                                          weights_file=weights_file)
            return coords.cpu(), confs.cpu()
        return fold_many(list(zip(input_files, templates)), one, gather=gather)

    dev_index = _compute_device(torch.device(device))
    key = (os.path.realpath(weights_file) if weights_file else '<default>', dev_index, int(streams))
    pool = _POOLS.get(key)
    if pool is None:
        pool = _POOLS[key] = StreamPool(load_weights(weights_file), dev_index, streams=int(streams))

    def batch(jobs):
        msas, tms = [], []
        for path, tmpl in jobs:
            alnmat = encode_aln(read_aln(path))
            if np.any(alnmat > 21):
                raise ValueError('alignment contains characters outside the residue alphabet')
            t = None
            if tmpl is not None:
                t = read_template(tmpl)
                if t.shape[0] != alnmat.shape[1]:
                    raise RuntimeError(f'Sizes of tensors must match: template has {t.shape[0]} CA atoms, '
                                       f'alignment has {alnmat.shape[1]} columns')
                t = torch.from_numpy(t)
            msas.append(torch.from_numpy(np.ascontiguousarray(alnmat)))
            tms.append(t)
        with torch.cuda.device(dev_index):
            out = pool.fold_all(msas, tms, max(iterations, 0), max(minsteps, 0))
            torch.cuda.synchronize()
        return [(c.cpu(), f.cpu()) for c, f in out]
    return fold_many_batched(list(zip(input_files, templates)), batch, gather=gather)


def confidence_summary(confs, threshold: float = 0.5) -> Dict[str, float]:
    """Convenience summary of the per-residue confidences aln_to_coords returns (the reference prints only their mean,
    `REMARK  CONF:`, predict.py:196): mean, min, max and the fraction of residues at or above `threshold`."""
    c = confs.detach().float().cpu().numpy() if torch.is_tensor(confs) else np.asarray(confs, dtype=np.float32)
    return {'mean': float(c.mean()), 'min': float(c.min()), 'max': float(c.max()),
            'fraction_confident': float((c >= threshold).mean()), 'threshold': float(threshold), 'residues': int(c.size)}


_RNAMES = {0: 'ALA', 1: 'ARG', 2: 'ASN', 3: 'ASP', 4: 'CYS', 5: 'GLN', 6: 'GLU', 7: 'GLY', 8: 'HIS', 9: 'ILE', 10: 'LEU',
           11: 'LYS', 12: 'MET', 13: 'PHE', 14: 'PRO', 15: 'SER', 16: 'THR', 17: 'TRP', 18: 'TYR', 19: 'VAL'}


def format_pdb(coords, confs, alnmat) -> str:
    """predict.py:195-208 -- byte-for-byte the reference's stdout (one D2H copy instead of 15*L .item() reads)."""
    c = coords.detach().cpu().numpy() if torch.is_tensor(coords) else np.asarray(coords)
    mean_conf = confs.mean().item() if torch.is_tensor(confs) else float(np.mean(confs))
    f = confs.detach().cpu() if torch.is_tensor(confs) else torch.as_tensor(confs)
    lines = [' '.join(('REMARK  CONF: ', str(mean_conf)))]      # == print("REMARK  CONF: ", x)
    atoms = (' N  ', ' CA ', ' C  ', ' O  ', ' CB ')
    atomnum = 1
    for ri in range(c.shape[0]):
        for ai, an in enumerate(atoms):
            if alnmat[0, ri] != 7 or ai != 4:
                lines.append('ATOM   %4d %s %s  %4d    %8.3f%8.3f%8.3f  1.00%6.2f' % (
                    atomnum, an, _RNAMES[int(alnmat[0, ri])], ri + 1,
                    c[ri, ai, 0].item(), c[ri, ai, 1].item(), c[ri, ai, 2].item(), f[ri]))
                atomnum += 1
    lines.append('END')
    return '\n'.join(lines) + '\n'


def run_dmpfold(argv=None):
    parser = argparse.ArgumentParser(description=(
        'The DMPfold2 method for fast and accurate protein structure prediction (B200-native engine). '
        'Prints a PDB format model file.'))
    parser.add_argument('-i', '--input_file', type=str, required=True, help='input sequence alignment in aln format')
    parser.add_argument('-d', '--device', type=str, default=default_device, required=False, help='device to run on')
    parser.add_argument('-t', '--template', type=str, required=False, help='use a PDB file as a template')
    parser.add_argument('-n', '--iterations', type=int, default=default_iterations, required=False,
                        help='number of iteration cycles')
    parser.add_argument('-m', '--minsteps', type=int, default=default_minsteps, required=False,
                        help='number of minimization steps')
    parser.add_argument('-w', '--model_weights', type=str, required=False, help='use a custom set of model weights')
    args = parser.parse_args(argv)
    coords, confs, alnmat = aln_to_coords(args.input_file, device=args.device, template=args.template,
                                          iterations=args.iterations, minsteps=args.minsteps,
                                          weights_file=args.model_weights, return_alnmat=True)
    sys.stdout.write(format_pdb(coords, confs, alnmat))
