"""Host-side mirror of the reference's decode interface, on top of the C ABI (include/brotlig_b200.h).

Names and argument meaning follow the reference:
  DecompressedSize(src)                       /root/reference/inc/BrotligDecoder.h:32
  DecodeCPU(src, output=None, feedbackProc)   /root/reference/inc/BrotligDecoder.h:33  (runs on the GPU here)
  DecodeGPU(useWarpDevice, src, output=None)  /root/reference/sample/BrotligGPUDecoder.h:24
plus `BrotligDecoder`, the object form (one CUDA context, batch and device-resident entry points).
Everything decodes on the GPU through libbrotlig_b200.so; there is no CPU path to fall back to.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Sequence

import numpy as np

from ._native import PROGRESS_FN, BgxPlanInfo, BgxStream, cuda_lib

BROTLIG_OK = 0
BROTLIG_ABORTED = 1
BROTLIG_ERROR_CORRUPT_STREAM = 14
BROTLIG_ERROR_INCORRECT_STREAM_FORMAT = 15
BROTLIG_ERROR_GENERIC = 16


class BrotligError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        super().__init__(f"BROTLIG_ERROR {code}" + (f": {msg}" if msg else ""))
        self.code = code


def _u8(a) -> np.ndarray:
    if isinstance(a, np.ndarray):
        return np.ascontiguousarray(a.view(np.uint8).reshape(-1))
    return np.frombuffer(a, dtype=np.uint8)


def DecompressedSize(src) -> int:
    s = _u8(src)
    if s.size < 8:
        raise BrotligError(BROTLIG_ERROR_CORRUPT_STREAM, "stream shorter than its header")
    return int(cuda_lib().bgx_decompressed_size(s.ctypes.data))


class Plan:
    """A device-resident decode: streams and outputs already live in HBM (torch tensors or raw pointers)."""

    def __init__(self, dec: "BrotligDecoder", handle: ctypes.c_void_p, keepalive):
        self._dec = dec
        self._h = handle
        self._keep = keepalive
        info = BgxPlanInfo()
        cuda_lib().bgx_plan_get_info(handle, ctypes.byref(info))
        self.info = {f: int(getattr(info, f)) for f, _ in BgxPlanInfo._fields_}

    def launch(self, cuda_stream: int | None = None) -> None:
        rc = cuda_lib().bgx_plan_launch(self._dec._ctx, self._h, ctypes.c_void_p(cuda_stream or 0))
        if rc:
            raise BrotligError(rc, self._dec.last_error())

    def finish(self) -> int:
        """Synchronises; returns the number of pages that failed to decode (0 = all good)."""
        bad = ctypes.c_uint32(0)
        rc = cuda_lib().bgx_plan_finish(self._dec._ctx, self._h, ctypes.byref(bad))
        if rc and rc != BROTLIG_ERROR_CORRUPT_STREAM:
            raise BrotligError(rc, self._dec.last_error())
        return int(bad.value)

    def close(self) -> None:
        if self._h:
            cuda_lib().bgx_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BrotligDecoder:
    def __init__(self, device: int = -1):
        lib = cuda_lib()
        ctx = ctypes.c_void_p()
        rc = lib.bgx_create(ctypes.byref(ctx), device)
        if rc:
            raise BrotligError(rc, "bgx_create failed: no usable CUDA device or kernel image (sm_100a)")
        self._ctx = ctx

    def last_error(self) -> str:
        return cuda_lib().bgx_last_error(self._ctx).decode()

    def close(self) -> None:
        if self._ctx:
            cuda_lib().bgx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host buffers in, host buffers out (what DecodeGPU / DecodeCPU bind)
    def decode_host(self, src, output: np.ndarray | None = None) -> tuple[np.ndarray, float]:
        outs, ms = self.decode_batch_host([src], None if output is None else [output])
        return outs[0], ms

    def decode_batch_host(self, srcs: Sequence, outputs: Sequence[np.ndarray] | None = None) -> tuple[list[np.ndarray], float]:
        lib = cuda_lib()
        ins = [_u8(s) for s in srcs]
        n = len(ins)
        if outputs is None:
            outputs = [np.empty(DecompressedSize(s), dtype=np.uint8) for s in ins]
        in_ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in ins])
        in_sizes = (ctypes.c_uint32 * n)(*[a.size for a in ins])
        out_ptrs = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outputs])
        out_sizes = (ctypes.c_uint32 * n)(*[o.size for o in outputs])
        ms = ctypes.c_double(0.0)
        rc = lib.bgx_decode_batch_host(self._ctx, n, in_ptrs, in_sizes, out_ptrs, out_sizes, ctypes.byref(ms))
        if rc:
            raise BrotligError(rc, self.last_error())
        return [o[: out_sizes[i]] for i, o in enumerate(outputs)], float(ms.value)

    def decode_host_progress(self, src, progress: Callable[[int, int], bool], output: np.ndarray | None = None,
                             pages_per_group: int = 0) -> tuple[np.ndarray, float]:
        """bgx_decode_host_progress: progress(page, num_pages) is called for every page after its group of pages was
        decoded; a true return stops the decode (the rest of the output is zero), as BrotligDecoder.cpp:318-325."""
        s = _u8(src)
        if output is None:
            output = np.empty(DecompressedSize(s), dtype=np.uint8)
        osz = ctypes.c_uint32(output.size)
        ms = ctypes.c_double(0.0)
        cb = PROGRESS_FN(lambda user, page, pages: 1 if progress(int(page), int(pages)) else 0)
        rc = cuda_lib().bgx_decode_host_progress(self._ctx, s.size, s.ctypes.data, ctypes.byref(osz), output.ctypes.data, ctypes.byref(ms),
                                                 cb, None, pages_per_group)
        if rc:
            raise BrotligError(rc, self.last_error())
        return output[: osz.value], float(ms.value)

    # ---- device-resident
    def plan(self, streams: Sequence[dict]) -> Plan:
        """streams: dicts with d_src (int device pointer), src_size, src_capacity, d_dst, dst_capacity,
        header (>= 16 bytes, host), optional page_begin / page_count."""
        n = len(streams)
        arr = (BgxStream * n)()
        for i, s in enumerate(streams):
            arr[i].d_src = s["d_src"]
            arr[i].src_size = s["src_size"]
            arr[i].src_capacity = s.get("src_capacity", s["src_size"])
            arr[i].d_dst = s["d_dst"]
            arr[i].dst_capacity = s["dst_capacity"]
            arr[i].page_begin = s.get("page_begin", 0)
            arr[i].page_count = s.get("page_count", 0)
            hdr = bytes(s["header"])[:16].ljust(16, b"\0")
            ctypes.memmove(arr[i].header, hdr, 16)
        handle = ctypes.c_void_p()
        rc = cuda_lib().bgx_plan_create(self._ctx, arr, n, ctypes.byref(handle))
        if rc:
            raise BrotligError(rc, self.last_error())
        return Plan(self, handle, arr)


def decode_batch_host_multi(decoders: Sequence[BrotligDecoder], srcs: Sequence, outputs: Sequence[np.ndarray] | None = None):
    """bgx_decode_batch_host_multi: one process, one decoder per device; whole streams are assigned to the decoders
    (balanced by compressed size) and decoded concurrently. Returns (outputs, kernel ms of the slowest device)."""
    lib = cuda_lib()
    ins = [_u8(s) for s in srcs]
    n = len(ins)
    if outputs is None:
        outputs = [np.empty(DecompressedSize(s), dtype=np.uint8) for s in ins]
    ctxs = (ctypes.c_void_p * len(decoders))(*[d._ctx for d in decoders])
    in_ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in ins])
    in_sizes = (ctypes.c_uint32 * n)(*[a.size for a in ins])
    out_ptrs = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outputs])
    out_sizes = (ctypes.c_uint32 * n)(*[o.size for o in outputs])
    ms = ctypes.c_double(0.0)
    rc = lib.bgx_decode_batch_host_multi(ctxs, len(decoders), n, in_ptrs, in_sizes, out_ptrs, out_sizes, ctypes.byref(ms))
    if rc:
        raise BrotligError(rc, "; ".join(d.last_error() for d in decoders))
    return [o[: out_sizes[i]] for i, o in enumerate(outputs)], float(ms.value)


_default: BrotligDecoder | None = None


def _decoder() -> BrotligDecoder:
    global _default
    if _default is None:
        _default = BrotligDecoder()
    return _default


def DecodeGPU(useWarpDevice: bool, src, output: np.ndarray | None = None) -> tuple[np.ndarray, float]:
    """Returns (decompressed bytes, kernel-only milliseconds). `useWarpDevice` is accepted and ignored."""
    del useWarpDevice
    return _decoder().decode_host(src, output)


def DecodeCPU(src, output: np.ndarray | None = None, feedbackProc: Callable[[int, str], bool] | None = None) -> np.ndarray:
    """Drop-in for BrotliG::DecodeCPU. Decodes on the GPU; feedbackProc(BROTLIG_PROGRESS, "100") is
    called once afterwards and may return True to report BROTLIG_ABORTED, as in the reference."""
    out, _ = _decoder().decode_host(src, output)
    if feedbackProc is not None and feedbackProc(0, "100.000000"):
        raise BrotligError(BROTLIG_ABORTED, "aborted by feedback callback")
    return out
