"""ctypes binding of libdmp2.so (include/dmp2.h) -- the only way the Python host reaches the GPU kernels.

There is no CPU fallback and no PyTorch implementation of any stage in this package: if the library is not
built (python -m dmpfold2_b200.build) or no sm_100 device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdmp2.so')

CONV_TC_F16X3, CONV_TC_F16, CONV_FFMA, CONV_TC_F16F8 = 0, 1, 2, 3
CONV_MODES = {'f16x3': CONV_TC_F16X3, 'f16': CONV_TC_F16, 'ffma': CONV_FFMA, 'f16f8': CONV_TC_F16F8}
STAGE_NAMES = ('vgru', 'hgru', 'msa_features_exposed', 'stem_base', 'recycling_passes', 'final_refine_backbone')

_lib = None

_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
_SIGS = {
    'dmp2_create': (_i, [C.POINTER(_vp), _i, _i, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(_i64)]),
    'dmp2_destroy': (None, [_vp]),
    'dmp2_last_error': (C.c_char_p, [_vp]),
    'dmp2_set_conv_mode': (_i, [_vp, _i]),
    'dmp2_launch_count': (_i64, [_vp]),
    'dmp2_stage_times': (_i, [_vp, C.POINTER(C.c_float), _i]),
    'dmp2_debug_eig_phases': (_i, [_vp, _i, C.POINTER(C.c_double), _i]),
    'dmp2_set_profile': (_i, [_vp, _i]),
    'dmp2_conv_profile': (_i, [_vp, C.POINTER(_i), C.POINTER(C.c_float)]),
    'dmp2_reserve': (_i, [_vp, _i, _i]),
    'dmp2_fold': (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    'dmp2_fold_host': (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp]),
    'dmp2_strip_rows': (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    'dmp2_strip_setup': (_i, [_vp, _i, _i, _i, _i, _vp, C.POINTER(_vp)]),
    'dmp2_strip_attach': (_i, [_vp, _vp]),
    'dmp2_strip_attach_local': (_i, [_vp, C.POINTER(_vp)]),
    'dmp2_strip_detach': (_i, [_vp]),
    'dmp2_fold_strip': (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    'dmp2_fold_strip_host': (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp]),
    'dmp2_reweight': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'dmp2_dca': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'dmp2_vgru': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'dmp2_hgru': (_i, [_vp, _vp, _i, _vp, _vp]),
    'dmp2_conv5_maxout': (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    'dmp2_resblock': (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    'dmp2_stem': (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    'dmp2_head': (_i, [_vp, _vp, _i, _vp, _vp]),
    'dmp2_resnet_pass': (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    'dmp2_head_mds': (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    'dmp2_eig_top8': (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    'dmp2_coord_gru': (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    'dmp2_refine': (_i, [_vp, _vp, _i, _i, _vp]),
    'dmp2_backbone': (_i, [_vp, _vp, _i, _vp, _vp]),
    'dmp2_gemm_tn_test': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    'dmp2_set_conv_sms': (_i, [_vp, _i]),
    'dmp2_set_graph': (_i, [_vp, _i]),
    'dmp2_set_conv_dynamic': (_i, [_vp, _i]),
    'dmp2_set_vgru_input': (_i, [_vp, _vp]),
}
EXPORTS = tuple(_SIGS)


def load_library() -> C.CDLL:
    """dlopen libdmp2.so and type every export of include/dmp2.h.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: build it with `python -m dmpfold2_b200.build` '
                               '(this engine has no CPU or PyTorch fallback)')
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class Dmp2Error(RuntimeError):
    pass


IPC_HANDLE_BYTES = 64
MAX_RANKS = 8


def strip_rows(l: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [r0, r1) of every L x L map owned by `rank` in a halo-sharded fold (dmp2_strip_rows; host arithmetic,
    works without a GPU).  Raises ValueError when L is too short for `world` strips."""
    r0, r1 = C.c_int(), C.c_int()
    if load_library().dmp2_strip_rows(int(l), int(world), int(rank), C.byref(r0), C.byref(r1)) != 0:
        raise ValueError(f'cannot split L={l} into {world} row strips (rank {rank}): every strip needs at least 2 rows')
    return r0.value, r1.value


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    """One engine per (device, stream): owns the repacked weights and the workspace on that device."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device_index: int = 0, conv_mode: Optional[str] = None):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise Dmp2Error('no CUDA device visible: the DMPfold2 B200 engine has no CPU fallback')
        self.device_index = int(device_index)
        self.device = torch.device('cuda', self.device_index)
        keep = [(k, v.detach().to('cpu', torch.float32).contiguous()) for k, v in state_dict.items()]
        n = len(keep)
        names = (C.c_char_p * n)(*[k.encode() for k, _ in keep])
        ptrs = (C.c_void_p * n)(*[v.data_ptr() for _, v in keep])
        numels = (C.c_int64 * n)(*[v.numel() for _, v in keep])
        handle = C.c_void_p()
        st = self.lib.dmp2_create(C.byref(handle), self.device_index, n, names, ptrs, numels)
        if st != 0:
            raise Dmp2Error(f'dmp2_create failed ({st}): {self.lib.dmp2_last_error(None).decode()}')
        self.h = handle
        if conv_mode is not None:
            self.set_conv_mode(conv_mode)

    def close(self):
        if getattr(self, 'h', None):
            self.lib.dmp2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _check(self, st: int, what: str):
        if st != 0:
            raise Dmp2Error(f'{what} failed ({st}): {self.lib.dmp2_last_error(self.h).decode()}')

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, a, dtype) -> torch.Tensor:
        t = torch.as_tensor(a)
        return t.to(self.device, dtype).contiguous()

    def _empty(self, *shape) -> torch.Tensor:
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def set_conv_mode(self, mode):
        m = CONV_MODES[mode] if isinstance(mode, str) else int(mode)
        self._check(self.lib.dmp2_set_conv_mode(self.h, m), 'dmp2_set_conv_mode')

    @property
    def launch_count(self) -> int:
        return int(self.lib.dmp2_launch_count(self.h))

    def stage_times(self) -> Dict[str, float]:
        buf = (C.c_float * 8)()
        n = self.lib.dmp2_stage_times(self.h, buf, 8)
        return {STAGE_NAMES[i]: float(buf[i]) for i in range(n)}

    def eig_phases(self, l: int):
        buf = (C.c_double * 4)()
        self._check(self.lib.dmp2_debug_eig_phases(self.h, int(l), buf, 4), 'dmp2_debug_eig_phases')
        return dict(zip(('tridiag_us', 'bisect_us', 'invit_us', 'backtransform_us'), list(buf)))

    def set_profile(self, on: bool):
        self._check(self.lib.dmp2_set_profile(self.h, 1 if on else 0), 'dmp2_set_profile')

    def reserve(self, l: int, n: int):
        """Size the workspace for alignments up to (n rows, l columns): later folds within those bounds allocate nothing
        (a fold that outgrows the workspace re-allocates, which synchronises the device once)."""
        with torch.cuda.device(self.device):
            self._check(self.lib.dmp2_reserve(self.h, int(l), int(n)), 'dmp2_reserve')

    def conv_profile(self) -> Tuple[int, float]:
        """(number of conv launches, their summed device time in ms) since the last call."""
        n, ms = C.c_int(), C.c_float()
        self._check(self.lib.dmp2_conv_profile(self.h, C.byref(n), C.byref(ms)), 'dmp2_conv_profile')
        return n.value, ms.value

    # ---- the hot path --------------------------------------------------------------------------
    def fold(self, msa: torch.Tensor, template_ca: Optional[torch.Tensor] = None, iterations: int = 10,
             minsteps: int = 100, vgru: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device tensors in, device tensors out; asynchronous on the current torch stream.
        msa uint8 (N,L) codes 0..21; template_ca float32 (L,3) or None -> coords (L,5,3), confs (L,).
        vgru: optional (L,512) float32 device tensor = self.vgru(msa), computed elsewhere (e.g. one scan over the
        columns of several alignments, parallel.StreamPool); the caller keeps it alive until the fold has run."""
        msa = self._dev(msa, torch.uint8)
        n, l = msa.shape
        if vgru is not None:
            if vgru.dtype != torch.float32 or tuple(vgru.shape) != (l, 512) or not vgru.is_contiguous() or vgru.device != self.device:
                raise ValueError('vgru must be a contiguous float32 (L, 512) tensor on the engine device')
            self._check(self.lib.dmp2_set_vgru_input(self.h, _ptr(vgru)), 'dmp2_set_vgru_input')
        tm = None
        if template_ca is not None:
            tm = self._dev(template_ca, torch.float32)
            if tuple(tm.shape) != (l, 3):
                raise ValueError(f'template has {tm.shape[0]} CA atoms, alignment has {l} columns')
        coords, conf = self._empty(l, 5, 3), self._empty(l)
        with torch.cuda.device(self.device):
            st = self.lib.dmp2_fold(self.h, _ptr(msa), n, l, _ptr(tm), int(iterations), int(minsteps), _ptr(coords),
                                    _ptr(conf), self._stream())
        self._check(st, 'dmp2_fold')
        return coords, conf

    def fold_host(self, msa: np.ndarray, template_ca: Optional[np.ndarray] = None, iterations: int = 10,
                  minsteps: int = 100) -> Tuple[np.ndarray, np.ndarray]:
        """Host buffers in/out through dmp2_fold_host (H2D + fold + D2H + sync) -- the end-to-end call."""
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        n, l = msa.shape
        tm = None if template_ca is None else np.ascontiguousarray(template_ca, dtype=np.float32)
        coords = np.empty((l, 5, 3), dtype=np.float32)
        conf = np.empty((l,), dtype=np.float32)
        st = self.lib.dmp2_fold_host(self.h, msa.ctypes.data_as(C.c_void_p), n, l,
                                     None if tm is None else tm.ctypes.data_as(C.c_void_p), int(iterations), int(minsteps),
                                     coords.ctypes.data_as(C.c_void_p), conf.ctypes.data_as(C.c_void_p))
        self._check(st, 'dmp2_fold_host')
        return coords, conf

    # ---- halo-sharded fold of one target over several GPUs ------------------------------------------
    def strip_setup(self, rank: int, world: int, l: int, reserve_n: int = 0) -> Tuple[bytes, int]:
        """Allocate this rank's exchange window for targets of length l (and the workspace for up to reserve_n
        alignment rows) -> (CUDA IPC handle, device address)."""
        handle = (C.c_ubyte * IPC_HANDLE_BYTES)()
        win = C.c_void_p()
        self._check(self.lib.dmp2_strip_setup(self.h, int(rank), int(world), int(l), int(reserve_n), C.cast(handle, C.c_void_p),
                                              C.byref(win)), 'dmp2_strip_setup')
        return bytes(handle), int(win.value)

    def strip_attach(self, handles) -> None:
        """handles: every rank's IPC handle in rank order (ranks in separate processes)."""
        blob = b''.join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self.lib.dmp2_strip_attach(self.h, C.cast(buf, C.c_void_p)), 'dmp2_strip_attach')

    def strip_attach_local(self, windows) -> None:
        """windows: every rank's window address in rank order (all ranks are engines of THIS process)."""
        arr = (C.c_void_p * len(windows))(*[C.c_void_p(w) for w in windows])
        self._check(self.lib.dmp2_strip_attach_local(self.h, arr), 'dmp2_strip_attach_local')

    def strip_detach(self) -> None:
        self._check(self.lib.dmp2_strip_detach(self.h), 'dmp2_strip_detach')

    def fold_strip(self, msa: torch.Tensor, template_ca: Optional[torch.Tensor] = None, iterations: int = 10,
                   minsteps: int = 100, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Collective form of fold(): every rank of the strip group calls it with the same arguments.
        out = preallocated (coords (L,5,3), confs (L,)) float32 tensors on the engine's device, optional."""
        msa = self._dev(msa, torch.uint8)
        n, l = msa.shape
        tm = None if template_ca is None else self._dev(template_ca, torch.float32)
        if tm is not None and tuple(tm.shape) != (l, 3):
            raise ValueError(f'template has {tm.shape[0]} CA atoms, alignment has {l} columns')
        coords, conf = out if out is not None else (self._empty(l, 5, 3), self._empty(l))
        with torch.cuda.device(self.device):
            st = self.lib.dmp2_fold_strip(self.h, _ptr(msa), n, l, _ptr(tm), int(iterations), int(minsteps), _ptr(coords),
                                          _ptr(conf), self._stream())
        self._check(st, 'dmp2_fold_strip')
        return coords, conf

    def fold_strip_host(self, msa: np.ndarray, template_ca: Optional[np.ndarray] = None, iterations: int = 10,
                        minsteps: int = 100) -> Tuple[np.ndarray, np.ndarray]:
        """Collective form of fold_host()."""
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        n, l = msa.shape
        tm = None if template_ca is None else np.ascontiguousarray(template_ca, dtype=np.float32)
        coords = np.empty((l, 5, 3), dtype=np.float32)
        conf = np.empty((l,), dtype=np.float32)
        st = self.lib.dmp2_fold_strip_host(self.h, msa.ctypes.data_as(C.c_void_p), n, l,
                                           None if tm is None else tm.ctypes.data_as(C.c_void_p), int(iterations), int(minsteps),
                                           coords.ctypes.data_as(C.c_void_p), conf.ctypes.data_as(C.c_void_p))
        self._check(st, 'dmp2_fold_strip_host')
        return coords, conf

    # ---- stage entry points (parity tests) -------------------------------------------------------
    def reweight(self, msa) -> torch.Tensor:
        msa = self._dev(msa, torch.uint8)
        n, l = msa.shape
        out = self._empty(n)
        self._check(self.lib.dmp2_reweight(self.h, _ptr(msa), n, l, _ptr(out), self._stream()), 'dmp2_reweight')
        return out

    def dca(self, msa) -> torch.Tensor:
        msa = self._dev(msa, torch.uint8)
        n, l = msa.shape
        out = self._empty(l, l, 442)
        self._check(self.lib.dmp2_dca(self.h, _ptr(msa), n, l, _ptr(out), self._stream()), 'dmp2_dca')
        return out

    def vgru(self, msa) -> torch.Tensor:
        msa = self._dev(msa, torch.uint8)
        n, l = msa.shape
        out = self._empty(l, 512)
        self._check(self.lib.dmp2_vgru(self.h, _ptr(msa), n, l, _ptr(out), self._stream()), 'dmp2_vgru')
        return out

    def hgru(self, x) -> torch.Tensor:
        x = self._dev(x, torch.float32)
        out = self._empty(x.shape[0], 512)
        self._check(self.lib.dmp2_hgru(self.h, _ptr(x), x.shape[0], _ptr(out), self._stream()), 'dmp2_hgru')
        return out

    def conv5_maxout(self, block: int, x_nhwc) -> torch.Tensor:
        x = self._dev(x_nhwc, torch.float32)
        l = x.shape[0]
        out = self._empty(l, l, 128)
        self._check(self.lib.dmp2_conv5_maxout(self.h, block, _ptr(x), l, _ptr(out), self._stream()), 'dmp2_conv5_maxout')
        return out

    def resblock(self, block: int, x_nhwc) -> torch.Tensor:
        x = self._dev(x_nhwc, torch.float32)
        l = x.shape[0]
        out = self._empty(l, l, 128)
        self._check(self.lib.dmp2_resblock(self.h, block, _ptr(x), l, _ptr(out), self._stream()), 'dmp2_resblock')
        return out

    def stem(self, mat1d_t, feat, dmap) -> torch.Tensor:
        """network.py:194 on the never-materialised 955-channel input of network.py:227-229 -> (L, L, 128) NHWC."""
        m = self._dev(mat1d_t, torch.float32)
        f = self._dev(feat, torch.float32)
        d = self._dev(dmap, torch.float32)
        l = m.shape[0]
        out = self._empty(l, l, 128)
        self._check(self.lib.dmp2_stem(self.h, _ptr(m), _ptr(f), _ptr(d), l, _ptr(out), self._stream()), 'dmp2_stem')
        return out

    def head(self, x_nhwc) -> torch.Tensor:
        """network.py:207 (resnet.17): (L, L, 128) NHWC -> (2, L, L)."""
        x = self._dev(x_nhwc, torch.float32)
        l = x.shape[0]
        out = self._empty(2, l, l)
        self._check(self.lib.dmp2_head(self.h, _ptr(x), l, _ptr(out), self._stream()), 'dmp2_head')
        return out

    def resnet_pass(self, mat1d_t, feat, dmap) -> torch.Tensor:
        m = self._dev(mat1d_t, torch.float32)
        f = self._dev(feat, torch.float32)
        d = self._dev(dmap, torch.float32)
        l = m.shape[0]
        out = self._empty(2, l, l)
        self._check(self.lib.dmp2_resnet_pass(self.h, _ptr(m), _ptr(f), _ptr(d), l, _ptr(out), self._stream()),
                    'dmp2_resnet_pass')
        return out

    def head_mds(self, head):
        h = self._dev(head, torch.float32)
        l = h.shape[-1]
        conf, m, mds = self._empty(l), self._empty(l, l), self._empty(l, 8)
        self._check(self.lib.dmp2_head_mds(self.h, _ptr(h), l, _ptr(conf), _ptr(m), _ptr(mds), self._stream()), 'dmp2_head_mds')
        return conf, m, mds

    def eig_top8(self, m):
        m = self._dev(m, torch.float32)
        l = m.shape[0]
        vals, vecs = self._empty(8), self._empty(l, 8)
        self._check(self.lib.dmp2_eig_top8(self.h, _ptr(m), l, _ptr(vals), _ptr(vecs), self._stream()), 'dmp2_eig_top8')
        return vals, vecs

    def coord_gru(self, mat1d_t, mds) -> torch.Tensor:
        m = self._dev(mat1d_t, torch.float32)
        d = self._dev(mds, torch.float32)
        l = m.shape[0]
        out = self._empty(l, 3)
        self._check(self.lib.dmp2_coord_gru(self.h, _ptr(m), _ptr(d), l, _ptr(out), self._stream()), 'dmp2_coord_gru')
        return out

    def refine(self, ca, steps: int) -> torch.Tensor:
        c = self._dev(ca, torch.float32).clone()
        self._check(self.lib.dmp2_refine(self.h, _ptr(c), c.shape[0], int(steps), self._stream()), 'dmp2_refine')
        return c

    def backbone(self, ca) -> torch.Tensor:
        c = self._dev(ca, torch.float32)
        out = self._empty(c.shape[0], 5, 3)
        self._check(self.lib.dmp2_backbone(self.h, _ptr(c), c.shape[0], _ptr(out), self._stream()), 'dmp2_backbone')
        return out

    def gemm_tn_test(self, a, b, mode='f16x3', chunk_k: int = 0) -> torch.Tensor:
        """C = A B^T on the tensor-core GEMM core; chunk_k = length of one tcgen05 accumulation chain (0 = all of K)."""
        a = self._dev(a, torch.float32)
        b = self._dev(b, torch.float32)
        m, k = a.shape
        n = b.shape[0]
        out = self._empty(m, n)
        self._check(self.lib.dmp2_gemm_tn_test(self.h, _ptr(a), _ptr(b), m, n, k, CONV_MODES[mode], int(chunk_k), _ptr(out),
                                               self._stream()), 'dmp2_gemm_tn_test')
        return out

    def set_conv_sms(self, sms: int):
        self._check(self.lib.dmp2_set_conv_sms(self.h, int(sms)), 'dmp2_set_conv_sms')

    def set_graph(self, on: bool = True):
        """Replay the recycling iterations (network.py:264-306) from a CUDA graph captured once per (L, workspace,
        kernel configuration) instead of re-enqueueing their ~60 launches each time; bit-identical results."""
        self._check(self.lib.dmp2_set_graph(self.h, 1 if on else 0), 'dmp2_set_graph')

    def set_conv_dynamic(self, on: bool = True):
        """Dynamic unit schedule of the persistent conv kernel (several folds in flight on one GPU, see dmp2.h)."""
        self._check(self.lib.dmp2_set_conv_dynamic(self.h, 1 if on else 0), 'dmp2_set_conv_dynamic')
