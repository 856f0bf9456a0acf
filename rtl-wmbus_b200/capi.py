"""ctypes mirror of include/wmbus_b200.h (one wrapper class, no logic of its own)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libwmbus_b200.so"


class WmbOpts(C.Structure):
    _fields_ = [("decimation", C.c_uint32), ("accurate_atan", C.c_uint8), ("remove_dc", C.c_uint8),
                ("rla_enabled", C.c_uint8), ("t2_enabled", C.c_uint8), ("t1c1_enabled", C.c_uint8),
                ("s1_enabled", C.c_uint8), ("simultaneous", C.c_uint8), ("show_algorithm", C.c_uint8),
                ("chunk_samples", C.c_uint32), ("warmup_samples", C.c_uint32),
                ("max_batch_mib", C.c_uint32), ("manual_frames", C.c_uint32), ("reserved", C.c_uint32 * 2),
                ("carrier_25khz", C.c_int32 * 2), ("prefilter", C.c_uint32)]


class WmbFrame(C.Structure):
    _fields_ = [("sync_sample", C.c_uint64), ("ordinal", C.c_uint64), ("chain", C.c_uint8),
                ("algo", C.c_uint8), ("truncated", C.c_uint8), ("reserved", C.c_uint8),
                ("nbits", C.c_uint32), ("bits", C.POINTER(C.c_uint32))]


class WmbStats(C.Structure):
    _fields_ = [("input_samples", C.c_uint64), ("decimated_samples", C.c_uint64), ("batches", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("lanes_run", C.c_uint64), ("lanes_rerun", C.c_uint64),
                ("candidates", (C.c_uint64 * 2) * 2), ("lines", (C.c_uint64 * 2) * 2),
                ("lines_crc_ok", (C.c_uint64 * 2) * 2), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("demod_kernel_ms", C.c_double), ("bitsync_kernel_ms", C.c_double),
                ("batch_device_ms", C.c_double), ("rl_fallbacks", C.c_uint64), ("host_batch_ms", C.c_double),
                ("host_gather_ms", C.c_double), ("host_decode_ms", C.c_double), ("overflow_batches", C.c_uint64)]


def library_path() -> str:
    return os.path.join(PKG_DIR, LIB_NAME)


def build(verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a into the in-tree libwmbus_b200.so and the rtl_wmbus_b200 CLI."""
    subprocess.run(["make", "-s", "-C", os.path.join(PKG_DIR, "csrc")] + ([] if not verbose else ["V=1"]), check=True)
    return library_path()


def _bind(lib):
    lib.wmb_default_opts.argtypes = [C.POINTER(WmbOpts)]
    lib.wmb_abi_version.restype = C.c_int
    lib.wmb_last_error.restype = C.c_char_p
    lib.wmb_version_string.restype = C.c_char_p
    lib.wmb_create.argtypes = [C.POINTER(WmbOpts), C.c_int, C.POINTER(C.c_void_p)]
    lib.wmb_destroy.argtypes = [C.c_void_p]
    lib.wmb_reset.argtypes = [C.c_void_p]
    lib.wmb_host_alloc.argtypes = [C.c_size_t]
    lib.wmb_host_alloc.restype = C.c_void_p
    lib.wmb_host_free.argtypes = [C.c_void_p]
    lib.wmb_push.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.wmb_push_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.wmb_poll.argtypes = [C.c_void_p, C.POINTER(WmbFrame), C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
    lib.wmb_decode_frames.argtypes = [C.c_void_p, C.POINTER(WmbFrame), C.c_size_t]
    lib.wmb_take_lines.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
    lib.wmb_take_lines.restype = C.c_size_t
    lib.wmb_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t,
                                C.POINTER(C.c_size_t), C.c_int]
    lib.wmb_process.restype = C.c_long
    lib.wmb_process_device.argtypes = lib.wmb_process.argtypes
    lib.wmb_process_device.restype = C.c_long
    lib.wmb_get_stats.argtypes = [C.c_void_p, C.POINTER(WmbStats)]
    lib.wmb_debug_copy_stage.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.wmb_debug_copy_stage.restype = C.c_long
    lib.wmb_debug_copy_bits.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.wmb_debug_copy_bits.restype = C.c_long
    lib.wmb_debug_copy_events.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.wmb_debug_copy_events.restype = C.c_long
    lib.wmb_debug_arith.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.wmb_seek.argtypes = [C.c_void_p, C.c_uint64]
    lib.wmb_set_line_window.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    lib.wmb_boundary_state.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.wmb_boundary_state.restype = C.c_long
    lib.wmb_pending_before.argtypes = [C.c_void_p, C.c_uint64]
    lib.wmb_pending_before.restype = C.c_long
    return lib


EXPORTS = ["wmb_reset", "wmb_host_alloc", "wmb_host_free", "wmb_default_opts", "wmb_abi_version", "wmb_last_error", "wmb_version_string", "wmb_create",
           "wmb_destroy", "wmb_push", "wmb_push_device", "wmb_poll", "wmb_decode_frames", "wmb_take_lines",
           "wmb_process", "wmb_process_device", "wmb_get_stats", "wmb_debug_copy_stage", "wmb_debug_copy_bits", "wmb_debug_copy_events", "wmb_debug_arith",
           "wmb_seek", "wmb_set_line_window", "wmb_boundary_state", "wmb_pending_before"]


def load_library(path: str | None = None):
    """dlopen libwmbus_b200.so.  Raises when it is missing: there is no fallback."""
    path = path or library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with __graft_entry__.build() "
                           f"(nvcc, sm_100a); rtl-wmbus_b200 has no CPU fallback")
    return _bind(C.CDLL(path))


def opts_from_flags(lib, flags: str = "", **kw) -> WmbOpts:
    """Build wmb_opts from a reference-style flag string, e.g. '-d 3 -s -o -v'."""
    o = WmbOpts()
    lib.wmb_default_opts(C.byref(o))
    toks = flags.split()
    i = 0
    while i < len(toks):
        t = toks[i]
        if t == "-o": o.remove_dc = 1
        elif t == "-a": o.accurate_atan = 0
        elif t == "-s": o.simultaneous = 1
        elif t == "-v": o.show_algorithm = 1
        elif t == "-f": pass
        elif t == "-d": i += 1; o.decimation = int(toks[i])
        elif t == "-r": i += 1; o.rla_enabled = 0 if toks[i] == "0" else 1
        elif t == "-t": i += 1; o.t2_enabled = 0 if toks[i] == "0" else 1
        elif t == "-p":
            i += 1
            if toks[i] in ("T", "t"): o.t1c1_enabled = 0
            elif toks[i] in ("S", "s"): o.s1_enabled = 0
            else: raise ValueError(toks[i])
        else:
            raise ValueError(f"unknown flag {t}")
        i += 1
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class WmbusB200:
    """One decoding context (== one rtl_wmbus process) on one GPU."""

    def __init__(self, flags: str = "", device: int = 0, lib=None, **tuning):
        self.lib = lib or load_library()
        self.opts = opts_from_flags(self.lib, flags, **tuning)
        self._ctx = C.c_void_p()
        rc = self.lib.wmb_create(C.byref(self.opts), device, C.byref(self._ctx))
        if rc != 0:
            raise RuntimeError(f"wmb_create failed ({rc}): {self.lib.wmb_last_error().decode()}")
        self._out = C.create_string_buffer(1 << 22)

    def close(self):
        if self._ctx:
            self.lib.wmb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self): return self
    def __exit__(self, *a): self.close()
    def __del__(self):
        try: self.close()
        except Exception: pass

    def _check(self, rc):
        if rc < 0:
            raise RuntimeError(f"libwmbus_b200 error {rc}: {self.lib.wmb_last_error().decode()}")
        return rc

    def _lines(self, n):
        if n <= 0:
            return []
        txt = C.string_at(self._out, n).decode("ascii")      # (.raw would copy the whole 4 MiB buffer)
        return txt.split("\n")[:-1] if txt.endswith("\n") else [l for l in txt.split("\n") if l]

    def _drain(self, taken, timestamp_mode):
        """wmb_process* hands out only the lines that fit the buffer; the rest stay queued -- fetch them too"""
        return self.take_lines(timestamp_mode) if taken else []

    def process(self, host_ptr, nbytes, flush=True, timestamp_mode=1, raw=False):
        """host_ptr: int address / ctypes pointer of cu8 bytes in host memory."""
        nl = C.c_size_t(0)
        n = self._check(self.lib.wmb_process(self._ctx, host_ptr, nbytes, int(flush), self._out,
                                             len(self._out), C.byref(nl), timestamp_mode))
        if raw:
            return self._raw(n, nl.value, timestamp_mode)
        return self._lines(n) + self._drain(nl.value, timestamp_mode)

    def process_bytes(self, data: bytes, flush=True, timestamp_mode=1):
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        return self.process(C.cast(buf, C.c_void_p), len(data), flush, timestamp_mode)

    def process_device(self, dev_ptr: int, nbytes: int, flush=True, timestamp_mode=1, raw=False):
        """raw=True: the text exactly as the C ABI hands it out (bytes, one line per datagram), not a list of str"""
        nl = C.c_size_t(0)
        n = self._check(self.lib.wmb_process_device(self._ctx, C.c_void_p(dev_ptr), nbytes, int(flush),
                                                    self._out, len(self._out), C.byref(nl), timestamp_mode))
        if raw:
            return self._raw(n, nl.value, timestamp_mode)
        return self._lines(n) + self._drain(nl.value, timestamp_mode)

    def _raw(self, n, taken, timestamp_mode):
        txt = C.string_at(self._out, n) if n > 0 else b""
        while taken:                                        # more lines than the buffer holds: fetch the rest
            k = C.c_size_t(0)
            m = self.lib.wmb_take_lines(self._ctx, self._out, len(self._out), C.byref(k), timestamp_mode)
            taken = k.value
            if taken:
                txt += C.string_at(self._out, m)
        return txt

    @staticmethod
    def split_lines(txt: bytes):
        return txt.decode("ascii").split("\n")[:-1] if txt else []

    def push(self, host_ptr, nbytes):
        self._check(self.lib.wmb_push(self._ctx, host_ptr, nbytes))

    def push_device(self, dev_ptr: int, nbytes: int):
        self._check(self.lib.wmb_push_device(self._ctx, dev_ptr, nbytes))

    def push_bytes(self, data: bytes):
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        self.push(C.cast(buf, C.c_void_p), len(data))

    def poll(self, flush=False, cap=1 << 16):
        arr = (WmbFrame * cap)()
        n = C.c_size_t(0)
        self._check(self.lib.wmb_poll(self._ctx, arr, cap, C.byref(n), int(flush)))
        return arr, n.value

    def poll_flush(self):
        """end of input: process what is buffered and finish the telegrams in flight (lines stay queued)"""
        n = C.c_size_t(0)
        self._check(self.lib.wmb_poll(self._ctx, None, 0, C.byref(n), 1))

    def decode_frames(self, arr, n):
        self._check(self.lib.wmb_decode_frames(self._ctx, arr, n))

    def take_lines(self, timestamp_mode=1):
        nl = C.c_size_t(0)
        out = []
        while True:
            n = self.lib.wmb_take_lines(self._ctx, self._out, len(self._out), C.byref(nl), timestamp_mode)
            if not nl.value:
                break
            out += self._lines(n)
        return out

    def reset(self):
        self._check(self.lib.wmb_reset(self._ctx))

    def seek(self, first_iq_sample: int):
        """reset + position the stream at an absolute IQ sample of the capture (time-chunk sharding)"""
        self._check(self.lib.wmb_seek(self._ctx, first_iq_sample))

    def set_line_window(self, sync_lo: int, sync_hi: int):
        self._check(self.lib.wmb_set_line_window(self._ctx, sync_lo, sync_hi))

    def pending_before(self, sync_hi: int) -> int:
        """telegrams in flight whose access-code match lies below decimated sample sync_hi"""
        return self._check(self.lib.wmb_pending_before(self._ctx, sync_hi))

    def boundary_state(self) -> bytes:
        if not hasattr(self, "_bbuf"):
            self._bbuf = C.create_string_buffer(1 << 22)        # (allocating and copying 4 MiB per call cost 2 ms)
        n = self._check(self.lib.wmb_boundary_state(self._ctx, self._bbuf, len(self._bbuf)))
        return C.string_at(self._bbuf, n)

    def stats(self) -> WmbStats:
        s = WmbStats()
        self._check(self.lib.wmb_get_stats(self._ctx, C.byref(s)))
        return s

    def debug_stage(self, chain: int, n: int):
        import numpy as np
        dphi = np.zeros(n, np.float32)
        rssi = np.zeros(n, np.uint8)
        got = self._check(self.lib.wmb_debug_copy_stage(self._ctx, chain, dphi.ctypes.data, rssi.ctypes.data, n))
        return dphi[:got], rssi[:got]

    def debug_bits(self, chain: int, which: int, n: int):
        """Unpacked 0/1 array of the last batch's data bits (0), time2 strobes (1) or clock signs (2)."""
        import numpy as np
        words = np.zeros((n + 31) // 32, np.uint32)
        got = self._check(self.lib.wmb_debug_copy_bits(self._ctx, chain, which, words.ctypes.data, len(words)))
        return np.unpackbits(words[:got].view(np.uint8), bitorder="little")[:n]

    def debug_events(self, chain: int, algo: int, cap: int = 1 << 22):
        """The last batch's bit events of one stream as (sample, rssi, reset, sync, bit) arrays."""
        import numpy as np
        ev = np.zeros(cap, np.uint64)
        got = self._check(self.lib.wmb_debug_copy_events(self._ctx, chain, algo, ev.ctypes.data, cap))
        ev = ev[:got]
        return dict(m=ev >> np.uint64(24), rssi=(ev >> np.uint64(16)) & np.uint64(0xFF), reset=(ev >> np.uint64(2)) & np.uint64(1),
                    sync=(ev >> np.uint64(1)) & np.uint64(1), bit=ev & np.uint64(1))

    def debug_arith(self, mode: int, y, x):
        """device arithmetic test hook: see wmb_debug_arith() in include/wmbus_b200.h"""
        import numpy as np
        y = np.ascontiguousarray(y, np.float32); x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(len(y), np.float32)
        self._check(self.lib.wmb_debug_arith(self._ctx, mode, y.ctypes.data, x.ctypes.data, out.ctypes.data, len(y)))
        return out
