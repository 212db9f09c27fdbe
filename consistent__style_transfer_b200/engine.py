"""WMDEngine: thin Python owner of a ``wmd_handle`` (include/wmd_b200.h).

numpy arrays go through the ``*_host`` entries (the library does its own chunked copies);
torch CUDA tensors go through the ``*_dev`` entries on the current torch stream with no host
synchronisation.  Nothing here computes: every number comes out of libwmd_b200.so.
"""
from __future__ import annotations

import ctypes
from itertools import chain
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import c_f32p, c_f64p, c_i32p, c_i64p

KERNEL_KINDS = ("nbow", "cost", "solve", "rwmd", "misc", "fused")
MODE_PYEMD = 0      # the reference's value: pyemd's 1e6-grid integer optimum (bit-faithful)
MODE_EXACT = 1      # additive: the real-valued transportation optimum in FP64
IDS_ARE_ROWS = 0x100  # flag: this call's ids are table rows although a token map is installed
_MODES = {"pyemd": MODE_PYEMD, "exact": MODE_EXACT, MODE_PYEMD: MODE_PYEMD, MODE_EXACT: MODE_EXACT}


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a: Optional[np.ndarray], ct):
    return None if a is None else a.ctypes.data_as(ct)


try:                                       # C packer (csrc/csrpack.c): ~5 ns per token instead of ~45 ns of interpreter work
    from . import build as _build
    _build.build_csrpack()
    from . import _csrpack
except Exception:                          # no C compiler / headers: the numpy route below does the same job
    _csrpack = None


def docs_to_csr(docs: Sequence[Sequence[int]], dtype=np.int32):
    """list of id lists -> (ids int32 [or int64], off int64); ids that do not fit become -1 (out of vocabulary)"""
    if _csrpack is not None and not isinstance(docs, np.ndarray):
        ids_b, off_b = _csrpack.pack(docs, 4 if dtype == np.int32 else 8)
        return np.frombuffer(ids_b, dtype=dtype), np.frombuffer(off_b, dtype=np.int64)
    off = np.zeros(len(docs) + 1, np.int64)
    if len(docs):
        np.cumsum(np.fromiter(map(len, docs), dtype=np.int64, count=len(docs)), out=off[1:])
    ids = np.fromiter(chain.from_iterable(docs), dtype=dtype, count=int(off[-1]))
    return ids, off


class WMDEngine:
    """One handle = one device copy of the embedding table + maps + workspace."""

    def __init__(self, vectors: np.ndarray, normalize: bool = False, device: int = 0,
                 rank: Optional[np.ndarray] = None, token_map: Optional[np.ndarray] = None,
                 distance_table: Optional[bool] = None):
        """distance_table: None = the library's policy (word-distance table when V * V * 4 bytes fit its budget),
        True / False = force the table / the direct path (``set_distance_table``)."""
        self._L = _lib.load()
        vectors = np.asarray(vectors)
        if vectors.ndim != 2:
            raise ValueError("vectors must be [V, d]")
        v = _np(vectors, np.float32)
        self.V, self.d = int(v.shape[0]), int(v.shape[1])
        self.device = int(device)
        h = _lib.c_handle()
        _lib.check(self._L.wmd_create(_ptr(v, c_f32p), self.V, self.d, self.d, int(bool(normalize)), self.device,
                                      ctypes.byref(h)))
        self._h = h
        if rank is not None:
            self.set_rank(rank)
        if token_map is not None:
            self.set_token_map(token_map)
        if distance_table is not None:
            self.set_distance_table(bool(distance_table))

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.wmd_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise RuntimeError("WMDEngine is closed")
        return self._h

    # -- configuration ----------------------------------------------------------------------
    def set_token_map(self, id_to_row: Optional[np.ndarray]):
        if id_to_row is None:
            _lib.check(self._L.wmd_set_token_map(self._handle(), None, 0))
            return
        m = _np(id_to_row, np.int32)
        _lib.check(self._L.wmd_set_token_map(self._handle(), _ptr(m, c_i32p), m.shape[0]))

    def set_rank(self, rank: Optional[np.ndarray]):
        if rank is None:
            _lib.check(self._L.wmd_set_rank(self._handle(), None, 0))
            return
        r = _np(rank, np.int32)
        _lib.check(self._L.wmd_set_rank(self._handle(), _ptr(r, c_i32p), r.shape[0]))

    def table(self) -> np.ndarray:
        out = np.empty((self.V, self.d), np.float32)
        _lib.check(self._L.wmd_get_table(self._handle(), _ptr(out, c_f32p)))
        return out

    # -- scoring: host buffers --------------------------------------------------------------
    def wmd_pairs(self, ids1, off1, ids2, off2, out: Optional[np.ndarray] = None,
                  status: Optional[np.ndarray] = None, mode="pyemd", ids_are_rows: bool = False):
        """WMD of CSR-packed pairs held in host memory. Returns (float64[B], int32 status[B]).
        mode "pyemd" (default) is the reference's value; "exact" the un-quantised FP64 optimum.
        ids_are_rows: the ids are embedding rows even though a token map is installed."""
        ids1, ids2 = _np(ids1, np.int32), _np(ids2, np.int32)
        off1, off2 = _np(off1, np.int64), _np(off2, np.int64)
        B = off1.shape[0] - 1
        if off2.shape[0] - 1 != B:
            raise ValueError("both sides must hold the same number of documents")
        if out is None:
            out = np.empty(B, np.float64)
        if status is None:
            status = np.empty(B, np.int32)
        _lib.check(self._L.wmd_pairs_host(self._handle(), _ptr(ids1, c_i32p), _ptr(off1, c_i64p),
                                          _ptr(ids2, c_i32p), _ptr(off2, c_i64p), B,
                                          _MODES[mode] | (IDS_ARE_ROWS if ids_are_rows else 0),
                                          _ptr(out, c_f64p), _ptr(status, c_i32p)))
        return out, status

    def submit_pairs(self, ids1, off1, ids2, off2, mode="pyemd", ids_are_rows: bool = False) -> int:
        """Asynchronous ``wmd_pairs``: stages the documents, queues copies and kernels and returns at once;
        ``wait_pairs`` collects the result.  One job per engine may be in flight.  Returns the pair count."""
        ids1, ids2 = _np(ids1, np.int32), _np(ids2, np.int32)
        off1, off2 = _np(off1, np.int64), _np(off2, np.int64)
        B = off1.shape[0] - 1
        if off2.shape[0] - 1 != B:
            raise ValueError("both sides must hold the same number of documents")
        _lib.check(self._L.wmd_pairs_submit(self._handle(), _ptr(ids1, c_i32p), _ptr(off1, c_i64p), _ptr(ids2, c_i32p),
                                            _ptr(off2, c_i64p), B, _MODES[mode] | (IDS_ARE_ROWS if ids_are_rows else 0)))
        return B

    def wait_pairs(self, npairs: int, out: Optional[np.ndarray] = None, status: Optional[np.ndarray] = None):
        if out is None:
            out = np.empty(npairs, np.float64)
        if status is None:
            status = np.empty(npairs, np.int32)
        _lib.check(self._L.wmd_pairs_wait(self._handle(), _ptr(out, c_f64p), _ptr(status, c_i32p)))
        return out, status

    # -- multi-GPU: the score gather fused into the kernels (sharding.PeerScores drives these) --------
    def peer_alloc(self, nbytes: int):
        """A zero-filled device buffer the other ranks of the box can map -> (device pointer, CUDA IPC handle bytes)."""
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        _lib.check(self._L.wmd_peer_alloc(self._handle(), int(nbytes), ctypes.byref(ptr), handle))
        return int(ptr.value), handle.raw

    def peer_open(self, handle: bytes) -> int:
        ptr = ctypes.c_void_p()
        _lib.check(self._L.wmd_peer_open(self._handle(), bytes(handle), ctypes.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr: int, opened: bool):
        _lib.check(self._L.wmd_peer_close(self._handle(), ctypes.c_void_p(ptr), int(bool(opened))))

    def set_fanout(self, out_ptrs: Sequence[int] = (), status_ptrs: Sequence[int] = ()):
        """Every score / status the pair entries store at pair index p from now on also goes to out_ptrs[k] + 8 p /
        status_ptrs[k] + 4 p (device pointers on this engine's device); no arguments: off."""
        n = len(out_ptrs)
        if n != len(status_ptrs):
            raise ValueError("one status array per score array")
        if n == 0:
            _lib.check(self._L.wmd_set_fanout(self._handle(), 0, None, None))
            return
        o = (ctypes.c_void_p * n)(*[int(x) for x in out_ptrs])
        t = (ctypes.c_void_p * n)(*[int(x) for x in status_ptrs])
        _lib.check(self._L.wmd_set_fanout(self._handle(), n, o, t))

    def workspace_bytes(self, npairs: int, max_len1: int, max_len2: int):
        """(upper estimate of the device bytes a pair call of that size needs, bytes the handle holds now)"""
        est = ctypes.c_int64(); res = ctypes.c_int64()
        _lib.check(self._L.wmd_workspace_bytes(self._handle(), int(npairs), int(max_len1), int(max_len2),
                                               ctypes.byref(est), ctypes.byref(res)))
        return int(est.value), int(res.value)

    def wmd_pairs_ptr(self, ids1_ptr: int, off1_ptr: int, ids2_ptr: int, off2_ptr: int, npairs: int,
                      out_ptr: int, status_ptr: int = 0):
        """Raw host-pointer form (pinned torch tensors' data_ptr()): no numpy wrapping, no copies."""
        f = self._L.wmd_pairs_host
        _lib.check(f(self._handle(), ctypes.cast(ids1_ptr, c_i32p), ctypes.cast(off1_ptr, c_i64p),
                     ctypes.cast(ids2_ptr, c_i32p), ctypes.cast(off2_ptr, c_i64p), int(npairs), MODE_PYEMD,
                     ctypes.cast(out_ptr, c_f64p), ctypes.cast(status_ptr, c_i32p) if status_ptr else None))

    def nbow(self, ids, off):
        """Per document: (rows int32, counts int32, weights float64) slices of length uniq[p] at off[p]."""
        ids, off = _np(ids, np.int32), _np(off, np.int64)
        n = off.shape[0] - 1
        rows = np.full(ids.shape[0], -1, np.int32)
        counts = np.zeros(ids.shape[0], np.int32)
        weights = np.zeros(ids.shape[0], np.float64)
        uniq = np.zeros(n, np.int32)
        _lib.check(self._L.wmd_nbow_host(self._handle(), _ptr(ids, c_i32p), _ptr(off, c_i64p), n,
                                         _ptr(rows, c_i32p), _ptr(counts, c_i32p), _ptr(weights, c_f64p),
                                         _ptr(uniq, c_i32p)))
        return rows, counts, weights, uniq

    def rwmd_pairs(self, ids1, off1, ids2, off2, want_argmin: bool = True):
        """Relaxed WMD lower bound. Returns dict(lb, l1, l2, argmin_rows, argmin_cols, status)."""
        ids1, ids2 = _np(ids1, np.int32), _np(ids2, np.int32)
        off1, off2 = _np(off1, np.int64), _np(off2, np.int64)
        B = off1.shape[0] - 1
        lb = np.empty(B, np.float64); l1 = np.empty(B, np.float64); l2 = np.empty(B, np.float64)
        st = np.empty(B, np.int32)
        am1 = np.full(ids1.shape[0], -1, np.int32) if want_argmin else None
        am2 = np.full(ids2.shape[0], -1, np.int32) if want_argmin else None
        _lib.check(self._L.wmd_rwmd_pairs_host(self._handle(), _ptr(ids1, c_i32p), _ptr(off1, c_i64p),
                                               _ptr(ids2, c_i32p), _ptr(off2, c_i64p), B,
                                               _ptr(lb, c_f64p), _ptr(l1, c_f64p), _ptr(l2, c_f64p),
                                               _ptr(am1, c_i32p), _ptr(am2, c_i32p), _ptr(st, c_i32p)))
        return dict(lb=lb, l1=l1, l2=l2, argmin_rows=am1, argmin_cols=am2, status=st)

    def emd_batch(self, P, Q, D, extra_mass_penalty: float = -1.0) -> np.ndarray:
        """pyemd.emd for a batch: P, Q float64 [B, n]; D float64 [n, n] (shared) or [B, n, n]."""
        P, Q, D = _np(P, np.float64), _np(Q, np.float64), _np(D, np.float64)
        if P.ndim != 2 or P.shape != Q.shape:
            raise ValueError("P and Q must both be [B, n]")
        B, n = P.shape
        shared = D.ndim == 2
        if D.shape != ((n, n) if shared else (B, n, n)):
            raise ValueError("D must be [n, n] or [B, n, n]")
        out = np.empty(B, np.float64)
        _lib.check(self._L.wmd_emd_batch_host(self._handle(), _ptr(P, c_f64p), _ptr(Q, c_f64p), _ptr(D, c_f64p), B, n,
                                              int(shared), float(extra_mass_penalty), _ptr(out, c_f64p)))
        return out

    def allpairs_topk(self, idsA, offA, idsB, offB, k: int, row_begin: int = 0, row_end: Optional[int] = None):
        """For rows [row_begin, row_end) of set A: the k documents of set B with the smallest WMD,
        ordered by (distance, index).  Returns (idx int32 [rows, k], dist float64 [rows, k], info)."""
        idsA, idsB = _np(idsA, np.int32), _np(idsB, np.int32)
        offA, offB = _np(offA, np.int64), _np(offB, np.int64)
        nA, nB = offA.shape[0] - 1, offB.shape[0] - 1
        row_end = nA if row_end is None else int(row_end)
        rows = row_end - int(row_begin)
        idx = np.empty((max(rows, 0), int(k)), np.int32)
        dist = np.empty((max(rows, 0), int(k)), np.float64)
        stats = (ctypes.c_int64 * 8)()
        ms = (ctypes.c_double * 4)()
        _lib.check(self._L.wmd_allpairs_topk_host(self._handle(), _ptr(idsA, c_i32p), _ptr(offA, c_i64p), nA,
                                                  _ptr(idsB, c_i32p), _ptr(offB, c_i64p), nB, int(k), int(row_begin), row_end,
                                                  _ptr(idx, c_i32p), _ptr(dist, c_f64p), stats, ms))
        info = {"bounds": int(stats[0]), "exact_round1": int(stats[1]), "exact_round2": int(stats[2]),
                "query_blocks": int(stats[3]), "ms_dist_table": ms[0], "ms_corpus_index": ms[1],
                "ms_bounds_select": ms[2], "ms_exact": ms[3]}
        return idx, dist, info

    def allpairs_topk_cuda(self, idsA, offA, idsB, offB, k: int, row_begin: int = 0, row_end: Optional[int] = None,
                           out_idx=None, out_dist=None):
        """``allpairs_topk`` with the result left in torch CUDA tensors (int32 [rows, k], float64 [rows, k]; passed in or
        allocated): what a multi-GPU job all-gathers over NCCL.  Returns info (and the tensors when it allocated them)."""
        import torch
        idsA, idsB = _np(idsA, np.int32), _np(idsB, np.int32)
        offA, offB = _np(offA, np.int64), _np(offB, np.int64)
        nA, nB = offA.shape[0] - 1, offB.shape[0] - 1
        row_end = nA if row_end is None else int(row_end)
        rows = max(row_end - int(row_begin), 0)
        dev = torch.device("cuda", self.device)
        made = out_idx is None or out_dist is None
        if out_idx is None:
            out_idx = torch.empty((rows, int(k)), dtype=torch.int32, device=dev)
        if out_dist is None:
            out_dist = torch.empty((rows, int(k)), dtype=torch.float64, device=dev)
        if not (out_idx.is_cuda and out_dist.is_cuda and out_idx.is_contiguous() and out_dist.is_contiguous() and
                out_idx.dtype == torch.int32 and out_dist.dtype == torch.float64 and
                out_idx.numel() == rows * int(k) and out_dist.numel() == rows * int(k)):
            raise ValueError("out_idx / out_dist must be contiguous CUDA tensors of [rows, k] int32 / float64")
        stats = (ctypes.c_int64 * 8)()
        ms = (ctypes.c_double * 4)()
        _lib.check(self._L.wmd_allpairs_topk_dev(self._handle(), _ptr(idsA, c_i32p), _ptr(offA, c_i64p), nA,
                                                 _ptr(idsB, c_i32p), _ptr(offB, c_i64p), nB, int(k), int(row_begin), row_end,
                                                 out_idx.data_ptr(), out_dist.data_ptr(), stats, ms))
        info = {"bounds": int(stats[0]), "exact_round1": int(stats[1]), "exact_round2": int(stats[2]),
                "query_blocks": int(stats[3]), "ms_dist_table": ms[0], "ms_corpus_index": ms[1],
                "ms_bounds_select": ms[2], "ms_exact": ms[3]}
        return (out_idx, out_dist, info) if made else info

    # -- scoring: device tensors (torch) ----------------------------------------------------
    def wmd_pairs_cuda(self, ids1, off1, ids2, off2, max_len1: int, max_len2: int, out=None, status=None, mode="pyemd"):
        """CSR pairs in torch CUDA tensors (int32 ids, int64 offsets) -> float64 CUDA tensor.
        Stream-ordered on torch's current stream; no host synchronisation."""
        import torch
        for t, dt in ((ids1, torch.int32), (ids2, torch.int32), (off1, torch.int64), (off2, torch.int64)):
            if not (t.is_cuda and t.dtype == dt and t.is_contiguous() and t.device.index == self.device):
                raise ValueError("expected contiguous CUDA tensors (int32 ids, int64 offsets) on the engine's device")
        B = off1.numel() - 1
        if out is None:
            out = torch.empty(B, dtype=torch.float64, device=ids1.device)
        if status is None:
            status = torch.empty(B, dtype=torch.int32, device=ids1.device)
        stream = torch.cuda.current_stream(ids1.device).cuda_stream
        _lib.check(self._L.wmd_pairs_dev(self._handle(), ids1.data_ptr(), off1.data_ptr(), ids1.numel(), int(max_len1),
                                         ids2.data_ptr(), off2.data_ptr(), ids2.numel(), int(max_len2),
                                         B, _MODES[mode], out.data_ptr(), status.data_ptr(), stream))
        return out, status

    def nbow_cuda(self, ids, off, max_len: int, want_weights: bool = True):
        """``nbow`` on torch CUDA tensors (int32 ids, int64 offsets); outputs stay on the device, stream-ordered."""
        import torch
        if not (ids.is_cuda and off.is_cuda and ids.dtype == torch.int32 and off.dtype == torch.int64 and
                ids.is_contiguous() and off.is_contiguous() and ids.device.index == self.device):
            raise ValueError("expected contiguous CUDA tensors (int32 ids, int64 offsets) on the engine's device")
        n = off.numel() - 1
        dev = ids.device
        rows = torch.full((ids.numel(),), -1, dtype=torch.int32, device=dev)
        counts = torch.zeros(ids.numel(), dtype=torch.int32, device=dev)
        weights = torch.zeros(ids.numel(), dtype=torch.float64, device=dev) if want_weights else None
        uniq = torch.zeros(n, dtype=torch.int32, device=dev)
        _lib.check(self._L.wmd_nbow_dev(self._handle(), ids.data_ptr(), off.data_ptr(), n, int(max_len), rows.data_ptr(),
                                        counts.data_ptr(), weights.data_ptr() if want_weights else None, uniq.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream))
        return rows, counts, weights, uniq

    def rwmd_pairs_cuda(self, ids1, off1, ids2, off2, max_len1: int, max_len2: int, want_argmin: bool = True):
        """``rwmd_pairs`` on torch CUDA tensors; every output stays on the device, stream-ordered, no host sync."""
        import torch
        for t, dt in ((ids1, torch.int32), (ids2, torch.int32), (off1, torch.int64), (off2, torch.int64)):
            if not (t.is_cuda and t.dtype == dt and t.is_contiguous() and t.device.index == self.device):
                raise ValueError("expected contiguous CUDA tensors (int32 ids, int64 offsets) on the engine's device")
        B = off1.numel() - 1
        dev = ids1.device
        f = lambda: torch.empty(B, dtype=torch.float64, device=dev)
        lb, l1, l2 = f(), f(), f()
        st = torch.empty(B, dtype=torch.int32, device=dev)
        am1 = torch.full((ids1.numel(),), -1, dtype=torch.int32, device=dev) if want_argmin else None
        am2 = torch.full((ids2.numel(),), -1, dtype=torch.int32, device=dev) if want_argmin else None
        _lib.check(self._L.wmd_rwmd_pairs_dev(self._handle(), ids1.data_ptr(), off1.data_ptr(), ids1.numel(), int(max_len1),
                                              ids2.data_ptr(), off2.data_ptr(), ids2.numel(), int(max_len2), B,
                                              lb.data_ptr(), l1.data_ptr(), l2.data_ptr(),
                                              am1.data_ptr() if want_argmin else None, am2.data_ptr() if want_argmin else None,
                                              st.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return dict(lb=lb, l1=l1, l2=l2, argmin_rows=am1, argmin_cols=am2, status=st)

    def wmd_pairs_padded(self, a, b, pad_id: int = 0, out=None, status=None):
        """Padded [B, L] torch CUDA id tensors (int32 or int64; pad_id skipped) -> float64[B] on device."""
        import torch
        if not (a.is_cuda and b.is_cuda and a.device.index == self.device and b.device.index == self.device):
            raise ValueError("expected CUDA tensors on the engine's device")
        if a.dtype != torch.int32:
            a = a.to(torch.int32)
        if b.dtype != torch.int32:
            b = b.to(torch.int32)
        a, b = a.contiguous(), b.contiguous()
        if not (a.is_cuda and b.is_cuda and a.dim() == 2 and b.dim() == 2 and a.shape[0] == b.shape[0]):
            raise ValueError("expected two [B, L] CUDA tensors")
        B = a.shape[0]
        if out is None:
            out = torch.empty(B, dtype=torch.float64, device=a.device)
        if status is None:
            status = torch.empty(B, dtype=torch.int32, device=a.device)
        stream = torch.cuda.current_stream(a.device).cuda_stream
        _lib.check(self._L.wmd_pairs_padded_dev(self._handle(), a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1],
                                                B, int(pad_id), MODE_PYEMD, out.data_ptr(), status.data_ptr(), stream))
        return out, status

    def wmd_pairs_torch(self, ids1, off1, ids2, off2, out=None, status=None):
        """Host CSR (numpy, pinned or pageable; offsets may start anywhere inside ``ids``) in, torch CUDA tensors
        (float64 scores, int32 status) out: the shape ``sharding.wmd_pairs_sharded`` wants as its ``score_fn`` --
        the library overlaps its chunked copies with the kernels and the scores stay on the device for the NCCL gather."""
        import torch
        dev = torch.device("cuda", self.device)
        ids1, ids2 = _np(ids1, np.int32), _np(ids2, np.int32)
        off1, off2 = _np(off1, np.int64), _np(off2, np.int64)
        B = off1.shape[0] - 1
        if out is None:
            out = torch.empty(B, dtype=torch.float64, device=dev)
        if status is None:
            status = torch.empty(B, dtype=torch.int32, device=dev)
        if B:
            _lib.check(self._L.wmd_pairs_host_in_dev_out(self._handle(), _ptr(ids1, c_i32p), _ptr(off1, c_i64p), _ptr(ids2, c_i32p),
                                                         _ptr(off2, c_i64p), B, MODE_PYEMD, out.data_ptr(), status.data_ptr()))
        return out, status

    # -- instrumentation --------------------------------------------------------------------
    def set_profiling(self, enabled: bool):
        _lib.check(self._L.wmd_set_profiling(self._handle(), int(bool(enabled))))

    def set_serial(self, enabled: bool):
        """One internal stream instead of two: per-kernel event times without co-scheduling effects."""
        _lib.check(self._L.wmd_set_serial(self._handle(), int(bool(enabled))))

    def set_distance_table(self, enabled: bool = True):
        """True: build (once) the V x V float32 word-distance table now and let the pair entries take their costs
        from it (the default whenever the table fits the library's budget; then the first scoring call builds it).
        False: the direct path that recomputes every distance from the embedding rows.  Results are bit-identical."""
        _lib.check(self._L.wmd_set_distance_table(self._handle(), int(bool(enabled))))

    def distance_table_info(self):
        b = ctypes.c_int64(); ms = ctypes.c_double(); en = ctypes.c_int32(); res = ctypes.c_int32()
        _lib.check(self._L.wmd_distance_table_info(self._handle(), ctypes.byref(b), ctypes.byref(ms), ctypes.byref(en),
                                                   ctypes.byref(res)))
        return {"bytes": int(b.value), "build_ms": float(ms.value), "enabled": bool(en.value), "resident": bool(res.value)}

    def profile(self, reset: bool = True):
        ms = (ctypes.c_double * len(KERNEL_KINDS))()
        n = (ctypes.c_int64 * len(KERNEL_KINDS))()
        _lib.check(self._L.wmd_get_profile(self._handle(), ms, n, int(reset)))
        return {k: {"ms": ms[i], "launches": int(n[i])} for i, k in enumerate(KERNEL_KINDS)}

    def last_stats(self):
        v = (ctypes.c_int64 * 6)()
        _lib.check(self._L.wmd_get_last_stats(self._handle(), v))
        keys = ("tokens", "uniques", "cells", "solved_pairs", "max_rows", "max_cols")
        return {k: int(v[i]) for i, k in enumerate(keys)}
