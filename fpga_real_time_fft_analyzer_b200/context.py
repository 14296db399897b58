"""FraContext - the thin Python face of libfra.so (ctypes; torch tensors as buffers).

torch is used for device memory and streams only; all arithmetic happens in the
hand-written CUDA kernels behind the C ABI (include/fra.h).  There is no fallback:
constructing a context without a CUDA device raises FraError."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._lib import FraError, lib

_WANT_ALL = ("filtered", "frames", "iq", "mag", "phase")


def _torch():
    import torch
    return torch


class FraContext:
    """One GPU's receive chain for `channels` independent channels.

    Mirrors the control surface of the reference's receiver backends
    (scripts/fft_analyzer_gui.py:567-613: send_command / send_filter_coefficients)
    through `command()`, and adds the batched data path."""

    def __init__(self, channels: int, fft_size: int = 16384, device: int = 0, flags: int = 0):
        self._L = lib()
        self._h = C.c_void_p()
        self.channels, self.n, self.device = int(channels), int(fft_size), int(device)
        self.flags = int(flags)
        self._pipe_keep = []              # FRA_PIPELINE: (x, out) of calls whose kernels may still be running
        self._async_calls = 0
        self._keep_alive = None
        rc = self._L.fra_create(C.byref(self._h), self.device, self.channels, self.n, flags)
        if rc != _abi.FRA_OK:
            self._h = C.c_void_p()
            raise FraError(rc, "fra_create")
        self._pinned = {}

    # ------------------------------------------------------------- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._L.fra_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, where):
        if rc != _abi.FRA_OK:
            raise FraError(rc, where, (self._L.fra_last_cuda_error(self._h) or b"").decode())

    # -------------------------------------------------------- control plane
    def command(self, data) -> bool:
        """Feed raw protocol bytes (an int is one byte).  Returns False while a
        0xF1 upload is still waiting for coefficient bytes."""
        if isinstance(data, int):
            data = bytes([data & 0xFF])
        data = bytes(data)
        rc = self._L.fra_command(self._h, data, len(data))
        if rc == _abi.FRA_ERR_BUSY:
            return False
        self._check(rc, "fra_command")
        return True

    def load_bank1(self, coeff12):
        arr = (C.c_int8 * 12)(*[int(v) for v in np.asarray(coeff12).reshape(12)])
        self._check(self._L.fra_load_bank1(self._h, arr), "fra_load_bank1")

    def load_sections(self, coeff6x6):
        """Six independent sections (int8 [6][6], RTL register order B0,B1,B2,A0,A1,A2): a
        superset of the 12-byte bank; becomes bank 1 until the next 12-byte upload."""
        arr = (C.c_int8 * 36)(*[int(v) for v in np.asarray(coeff6x6).reshape(36)])
        self._check(self._L.fra_load_sections(self._h, arr), "fra_load_sections")

    def sections(self):
        """int8 [6][6]: the coefficients the current mode filters with."""
        arr = (C.c_int8 * 36)()
        self._check(self._L.fra_get_sections(self._h, arr), "fra_get_sections")
        return np.array(arr[:], dtype=np.int8).reshape(6, 6)

    def set_mag_average(self, alpha: float):
        """alpha = 1: plain magnitudes; 0 < alpha < 1: the `mag` output buffer (pass the same one every
        call via out=) holds the exponential moving average mag += alpha * (|bin| - mag)."""
        self._check(self._L.fra_set_mag_average(self._h, C.c_float(alpha)), "fra_set_mag_average")

    def set_mode(self, mode: int):
        self._check(self._L.fra_set_mode(self._h, mode), "fra_set_mode")

    def reset(self):
        self._check(self._L.fra_reset(self._h), "fra_reset")

    @property
    def mode(self) -> int:
        v = C.c_uint8()
        self._check(self._L.fra_get_mode(self._h, C.byref(v)), "fra_get_mode")
        return v.value

    @property
    def transport(self) -> int:
        v = C.c_uint8()
        self._check(self._L.fra_get_transport(self._h, C.byref(v)), "fra_get_transport")
        return v.value

    def bank(self, index: int) -> np.ndarray:
        arr = (C.c_int8 * 12)()
        self._check(self._L.fra_get_bank(self._h, index, arr), "fra_get_bank")
        return np.array(list(arr), dtype=np.int8)

    def counters(self) -> dict:
        v = [C.c_uint64() for _ in range(4)]
        self._check(self._L.fra_get_counters(self._h, *[C.byref(x) for x in v]), "fra_get_counters")
        return dict(zip(("start", "request", "reset", "upload"), (x.value for x in v)))

    @staticmethod
    def window_rom() -> np.ndarray:
        arr = (C.c_int16 * _abi.FRA_WINDOW_LEN)()
        rc = lib().fra_window_rom(arr)
        if rc != _abi.FRA_OK:
            raise FraError(rc, "fra_window_rom")
        return np.frombuffer(arr, dtype=np.int16).copy()

    # ------------------------------------------------------------ data path
    def _alloc_outputs(self, want, device):
        torch = _torch()
        c, n = self.channels, self.n
        shapes = {"filtered": ((c, n), torch.int16), "frames": ((c, 4 * n), torch.uint8),
                  "iq": ((c, n, 2), torch.float32), "mag": ((c, n), torch.float32),
                  "phase": ((c, n), torch.float32)}
        return {k: torch.empty(shapes[k][0], dtype=shapes[k][1], device=device) for k in want}

    @staticmethod
    def _outputs_struct(out):
        p = lambda k: out[k].data_ptr() if k in out else None
        return _abi.FraOutputs(p("filtered"), p("frames"), p("iq"), p("mag"), p("phase"))

    def process(self, x, continuous=False, log2_scale=None, want=("frames",), out=None):
        """One step on device tensors.  x: int16 [C, N] on this context's GPU.
        Enqueued on torch's current stream; returns {name: tensor}.

        FRA_PIPELINE contexts: the kernels run on the library's two internal streams (they only
        wait for the current stream at the time of the call) and the FFT of this call is enqueued
        by the NEXT call, join() or sync().  The returned tensors are complete only after join()
        (device-side) or sync() (host-side); until then the context keeps x and the outputs alive,
        so that torch's caching allocator cannot hand their memory to someone else."""
        torch = _torch()
        if not (x.is_cuda and x.dtype == torch.int16 and x.is_contiguous()
                and x.numel() == self.channels * self.n and x.device.index == self.device):
            raise ValueError("x must be a contiguous int16 CUDA tensor of shape [channels, fft_size] on this device")
        if out is None:
            out = self._alloc_outputs(want, x.device)
        o = self._outputs_struct(out)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        ls = _abi.FRA_SCALE_DEFAULT if log2_scale is None else int(log2_scale)
        self._check(self._L.fra_process(self._h, x.data_ptr(), int(bool(continuous)), ls, C.byref(o),
                                        C.c_void_p(stream)), "fra_process")
        if self.flags & _abi.FRA_PIPELINE:
            self._pipe_keep.append((x, out))
        return out

    def pinned(self, name, shape, dtype):
        """A reusable pinned host tensor (the e2e path copies from / into these)."""
        torch = _torch()
        key = (name, tuple(shape), dtype)
        if key not in self._pinned:
            self._pinned[key] = torch.empty(shape, dtype=dtype, pin_memory=True)
        return self._pinned[key]

    def process_host(self, x, continuous=False, log2_scale=None, want=("frames",)):
        """One step on HOST buffers (the GpuReceiver call): H2D, kernels, D2H inside.
        x: numpy int16 [C, N] or a CPU torch tensor (pinned for full speed).
        Returns {name: CPU torch tensor (pinned, reused between calls)}."""
        torch = _torch()
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.int16))
        if x.is_cuda or x.dtype != torch.int16 or not x.is_contiguous() or x.numel() != self.channels * self.n:
            raise ValueError("x must be a contiguous int16 host array of shape [channels, fft_size]")
        c, n = self.channels, self.n
        shapes = {"filtered": ((c, n), torch.int16), "frames": ((c, 4 * n), torch.uint8),
                  "iq": ((c, n, 2), torch.float32), "mag": ((c, n), torch.float32),
                  "phase": ((c, n), torch.float32)}
        out = {k: self.pinned(k, *shapes[k]) for k in want}
        o = self._outputs_struct(out)
        ls = _abi.FRA_SCALE_DEFAULT if log2_scale is None else int(log2_scale)
        self._check(self._L.fra_process_host(self._h, x.data_ptr(), int(bool(continuous)), ls, C.byref(o)),
                    "fra_process_host")
        return out

    def process_host_async(self, x, continuous=False, log2_scale=None, want=("frames",)):
        """process_host without the final wait: returns (outputs, ticket); the outputs (pinned
        host tensors, two alternating sets) are valid after host_wait(ticket).  x must stay
        untouched until then.  Two calls may be in flight, so frame i+1 uploads while frame i
        downloads."""
        torch = _torch()
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.int16))
        if x.is_cuda or x.dtype != torch.int16 or not x.is_contiguous() or x.numel() != self.channels * self.n:
            raise ValueError("x must be a contiguous int16 host array of shape [channels, fft_size]")
        c, n = self.channels, self.n
        shapes = {"filtered": ((c, n), torch.int16), "frames": ((c, 4 * n), torch.uint8),
                  "iq": ((c, n, 2), torch.float32), "mag": ((c, n), torch.float32),
                  "phase": ((c, n), torch.float32)}
        slot = self._async_calls & 1
        self._async_calls += 1
        out = {k: self.pinned(f"{k}#{slot}", *shapes[k]) for k in want}
        o = self._outputs_struct(out)
        ls = _abi.FRA_SCALE_DEFAULT if log2_scale is None else int(log2_scale)
        ticket = C.c_uint64(0)
        self._keep_alive = (x, self._keep_alive[0] if getattr(self, "_keep_alive", None) else None)
        self._check(self._L.fra_process_host_async(self._h, x.data_ptr(), int(bool(continuous)), ls, C.byref(o),
                                                   C.byref(ticket)), "fra_process_host_async")
        return out, ticket.value

    def host_wait(self, ticket):
        self._check(self._L.fra_host_wait(self._h, C.c_uint64(int(ticket))), "fra_host_wait")

    def set_host_half_share(self, share=-1.0):
        """FRA_HOST_HALF_SPECTRUM: share of the frames sent as half spectra, 0..1; negative = adaptive (default)."""
        self._check(self._L.fra_set_host_half_share(self._h, C.c_double(float(share))), "fra_set_host_half_share")

    def host_transfer(self):
        """(h2d_bytes, d2h_bytes) of the last host call and the half-spectrum share the next one will use."""
        h2d, d2h, share = C.c_uint64(0), C.c_uint64(0), C.c_double(0.0)
        wait, mirror = C.c_double(0.0), C.c_double(0.0)
        self._check(self._L.fra_get_host_transfer(self._h, C.byref(h2d), C.byref(d2h), C.byref(share), C.byref(wait),
                                                  C.byref(mirror)), "fra_get_host_transfer")
        self.host_wait_s, self.host_mirror_s = wait.value, mirror.value
        return h2d.value, d2h.value, share.value

    def get_state(self):
        torch = _torch()
        st = torch.empty((self.channels, 6, 4), dtype=torch.int16, device=f"cuda:{self.device}")
        stream = torch.cuda.current_stream(st.device).cuda_stream
        self._check(self._L.fra_get_state(self._h, st.data_ptr(), C.c_void_p(stream)), "fra_get_state")
        return st

    def set_state(self, st):
        torch = _torch()
        st = st.to(device=f"cuda:{self.device}", dtype=torch.int16).contiguous()
        if st.numel() != self.channels * 24:
            raise ValueError("state must have shape [channels, 6, 4]")
        stream = torch.cuda.current_stream(st.device).cuda_stream
        self._check(self._L.fra_set_state(self._h, st.data_ptr(), C.c_void_p(stream)), "fra_set_state")
        torch.cuda.current_stream(st.device).synchronize()

    def iir_stream(self, x, continuous=False, exact=False):
        """Window + IIR12 of one long stream (device int16 [n]); returns (y, stats dict)."""
        torch = _torch()
        if not (x.is_cuda and x.dtype == torch.int16 and x.is_contiguous() and x.dim() == 1):
            raise ValueError("x must be a contiguous 1-D int16 CUDA tensor")
        y = torch.empty_like(x)
        st = _abi.FraStreamStats()
        torch.cuda.current_stream(x.device).synchronize()
        self._check(self._L.fra_iir_stream(self._h, x.data_ptr(), y.data_ptr(), x.numel(), int(bool(continuous)),
                                           int(bool(exact)), C.byref(st)), "fra_iir_stream")
        return y, {k: getattr(st, k) for k, _ in st._fields_}

    def fft_only(self, x):
        """FFT alone: int16 [B, N] on the GPU -> complex64 [B, N]."""
        torch = _torch()
        if not (x.is_cuda and x.dtype == torch.int16 and x.is_contiguous() and x.shape[-1] == self.n):
            raise ValueError("x must be a contiguous int16 CUDA tensor [batch, fft_size]")
        batch = x.numel() // self.n
        iq = torch.empty((batch, self.n, 2), dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        self._check(self._L.fra_fft_only(self._h, x.data_ptr(), batch, iq.data_ptr(), C.c_void_p(stream)),
                    "fra_fft_only")
        return torch.view_as_complex(iq)

    def profile(self, on: bool = True):
        """Record CUDA events around each kernel of process() (bench.py's roofline)."""
        self._check(self._L.fra_profile_enable(self._h, int(bool(on))), "fra_profile_enable")

    def profile_last(self):
        """(ms window+IIR kernel, ms FFT+pack kernel) of the last process() call."""
        a, b = C.c_float(), C.c_float()
        self._check(self._L.fra_profile_last(self._h, C.byref(a), C.byref(b)), "fra_profile_last")
        return a.value, b.value

    def sync(self):
        """Host waits for everything the context has enqueued."""
        self._check(self._L.fra_sync(self._h), "fra_sync")
        self._pipe_keep.clear()

    def join(self):
        """FRA_PIPELINE contexts: torch's current stream waits (on the device) for all work
        enqueued by earlier process() calls; outputs may be read on that stream afterwards."""
        import torch
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._L.fra_join(self._h, C.c_void_p(stream)), "fra_join")
        # the current stream is now ordered behind every internal kernel: a tensor released here
        # is recycled by the caching allocator in stream order, i.e. after those kernels
        self._pipe_keep.clear()

    @property
    def last_kernel_count(self) -> int:
        return self._L.fra_last_kernel_count(self._h)
