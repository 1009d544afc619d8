"""Runs the UNMODIFIED reference (staged in oracle/_ref by oracle/make_ref.py) through its own public API
`dmpfold.aln_to_coords` (predict.py:74-158) on the host CPU.  TEST / BASELINE INFRASTRUCTURE ONLY: imported by
`bench.py --impl reference`, bench.py's `cpu_baseline` leg and tools/; never by the product.

The one modification the reference needs on torch >= 1.13 is installed here before it is imported: `torch.symeig`
was removed (network.py:247, :292).  The replacement is "oracle B" of SURVEY.md A.1 -- eigh on the upper triangle
plus the canonical eigenvector sign that makes the reference reproducible across thread counts.
"""
import os
import sys
import tempfile
from typing import Optional, Tuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
WEIGHTS_DIR = os.path.join(HERE, '..', 'dmpfold2_b200', 'trained_model')
_LETTERS = 'ARNDCQEGHILKMFPSTWYVX-'          # code -> letter (20 = unknown class, 21 = gap), predict.py:124-128
_merged: Optional[str] = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, 'dmpfold', 'network.py'))


def _install_shim():
    def _symeig(a, eigenvectors=False, upper=True):
        w, v = torch.linalg.eigh(a, UPLO='U' if upper else 'L')
        idx = v.abs().argmax(dim=-2, keepdim=True)
        return w, v * torch.gather(v, -2, idx).sign()
    torch.symeig = _symeig


def load():
    """Import the staged reference package (with the symeig shim)."""
    if not available():
        raise RuntimeError('oracle/_ref is missing: run `python oracle/make_ref.py` in the build container')
    _install_shim()
    # loaded under its own module name: the repository also ships a `dmpfold` alias package for the B200 engine
    import importlib.util
    if 'dmpfold_reference' in sys.modules:
        return sys.modules['dmpfold_reference']
    pkg = os.path.join(REF_DIR, 'dmpfold')
    spec = importlib.util.spec_from_file_location('dmpfold_reference', os.path.join(pkg, '__init__.py'),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules['dmpfold_reference'] = mod
    spec.loader.exec_module(mod)
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(REF_DIR)), mod.__file__
    return mod


def merged_weights_file() -> str:
    """The reference's `weights_file` argument takes ONE state_dict file (predict.py:93-95); merge the two staged parts
    once per process into a scratch file (untimed set-up)."""
    global _merged
    if _merged is None or not os.path.isfile(_merged):
        sd = torch.load(os.path.join(WEIGHTS_DIR, 'FINAL_fullmap_e2e_model_part1.pt'), map_location='cpu')
        sd.update(torch.load(os.path.join(WEIGHTS_DIR, 'FINAL_fullmap_e2e_model_part2.pt'), map_location='cpu'))
        fd, path = tempfile.mkstemp(suffix='.pt', prefix='dmpfold_ref_weights_')
        os.close(fd)
        torch.save(sd, path)
        _merged = path
    return _merged


def write_aln(path: str, msa: np.ndarray) -> None:
    """Encoded alignment (codes 0..21) -> the text the reference parses (inverse of predict.py:124-128)."""
    with open(path, 'w') as fh:
        for row in np.asarray(msa):
            fh.write(''.join(_LETTERS[int(c)] for c in row) + '\n')


def fold(msa: np.ndarray, iterations: int = 10, minsteps: int = 100, template: Optional[str] = None,
         threads: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """One call of the reference's aln_to_coords on the CPU.  Returns coords (L,5,3), confs (L,)."""
    ref = load()
    if threads:
        torch.set_num_threads(threads)
    with tempfile.TemporaryDirectory() as tmp:
        aln = os.path.join(tmp, 'target.aln')
        write_aln(aln, msa)
        coords, confs = ref.aln_to_coords(aln, device='cpu', template=template, iterations=iterations, minsteps=minsteps,
                                          weights_file=merged_weights_file())
    return coords.detach().numpy(), confs.detach().numpy()
