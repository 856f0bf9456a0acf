"""ctypes binding of the CPU oracle (oracle/_ref/liboracle.so) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")


class OrcOpts(C.Structure):
    _fields_ = [("decimation", C.c_uint32), ("accurate_atan", C.c_uint8), ("remove_dc", C.c_uint8),
                ("rla_enabled", C.c_uint8), ("t2_enabled", C.c_uint8), ("t1c1_enabled", C.c_uint8),
                ("s1_enabled", C.c_uint8), ("simultaneous", C.c_uint8), ("show_algorithm", C.c_uint8),
                ("real_timestamp", C.c_uint8), ("carrier_25khz", C.c_int32 * 2), ("prefilter", C.c_uint32)]


class OrcEvent(C.Structure):
    _fields_ = [("m", C.c_uint64), ("bit", C.c_uint8), ("sync", C.c_uint8), ("reset", C.c_uint8),
                ("rssi", C.c_uint8)]


EVENT_DTYPE = np.dtype([("m", "<u8"), ("bit", "u1"), ("sync", "u1"), ("reset", "u1"), ("rssi", "u1"),
                        ("pad", "u1", 4)])

_lib = None


def build():
    """(Re)build the oracle (and, where /root/reference exists, oracle/_ref)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
    if os.path.exists("/root/reference/rtl_wmbus.c"):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(REF_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
        u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
        L.orc_default_opts.argtypes = [C.POINTER(OrcOpts)]
        L.orc_atan2f.argtypes = [C.c_float, C.c_float]; L.orc_atan2f.restype = C.c_float
        L.orc_num_decimated.argtypes = [C.c_size_t, C.c_uint32]; L.orc_num_decimated.restype = C.c_size_t
        L.orc_frontend.argtypes = [u8p, C.c_size_t, C.POINTER(OrcOpts), C.c_int, f32p, f32p]
        L.orc_discriminator.argtypes = [f32p, f32p, C.c_size_t, C.c_int, f32p]
        L.orc_fir.argtypes = [f32p, C.c_size_t, C.c_int, f32p]
        L.orc_dcblock.argtypes = [f32p, C.c_size_t]
        L.orc_slicer.argtypes = [f32p, C.c_size_t, u8p]
        L.orc_rssi.argtypes = [f32p, f32p, C.c_size_t, f32p]
        L.orc_clock.argtypes = [f32p, C.c_size_t, C.c_int, u8p]
        L.orc_clock_state.argtypes = [f32p, C.c_size_t, C.c_int, f32p, u8p]
        L.orc_time2_strobe.argtypes = [u8p, C.c_size_t, u8p]
        L.orc_time2_events.argtypes = [u8p, u8p, f32p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_time2_events.restype = C.c_size_t
        L.orc_runlength_events.argtypes = [u8p, f32p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_runlength_events.restype = C.c_size_t
        L.orc_run.argtypes = [u8p, C.c_size_t, C.POINTER(OrcOpts), C.c_char_p, C.c_size_t,
                              C.POINTER(C.c_size_t)]
        L.orc_run.restype = C.c_size_t
        L.orc_frame_t1c1.argtypes = [u8p, u8p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_size_t,
                                     C.POINTER(C.c_int)]
        L.orc_frame_t1c1.restype = C.c_size_t
        L.orc_frame_s1.argtypes = L.orc_frame_t1c1.argtypes
        L.orc_frame_s1.restype = C.c_size_t
        L.orc_crc16.argtypes = [u8p, C.c_size_t]; L.orc_crc16.restype = C.c_uint16
        _lib = L
    return _lib


def opts(decimation=2, accurate_atan=1, remove_dc=0, rla=1, t2=1, t1c1=1, s1=1, simultaneous=0,
         show_algorithm=0):
    o = OrcOpts()
    lib().orc_default_opts(C.byref(o))
    o.decimation = decimation; o.accurate_atan = accurate_atan; o.remove_dc = remove_dc
    o.rla_enabled = rla; o.t2_enabled = t2; o.t1c1_enabled = t1c1; o.s1_enabled = s1
    o.simultaneous = simultaneous; o.show_algorithm = show_algorithm
    return o


def opts_from_flags(flags):
    """Parse a reference-style flag string ('-d 3 -s -o') into OrcOpts."""
    o = opts()
    toks = flags.split()
    i = 0
    while i < len(toks):
        t = toks[i]
        if t == "-o": o.remove_dc = 1
        elif t == "-a": o.accurate_atan = 0
        elif t == "-s": o.simultaneous = 1
        elif t == "-v": o.show_algorithm = 1
        elif t == "-d": i += 1; o.decimation = int(toks[i])
        elif t == "-r": i += 1; o.rla_enabled = 0 if toks[i] == "0" else 1
        elif t == "-t": i += 1; o.t2_enabled = 0 if toks[i] == "0" else 1
        elif t == "-p":
            i += 1
            if toks[i] in "Tt": o.t1c1_enabled = 0
            else: o.s1_enabled = 0
        else:
            raise ValueError(t)
        i += 1
    return o


def stages(cu8, o, chain):
    """Run every oracle stage for one chain; returns dict of arrays."""
    L = lib()
    cu8 = np.ascontiguousarray(cu8, dtype=np.uint8)
    n_iq = (len(cu8) - len(cu8) % 4096) // 2
    M = L.orc_num_decimated(n_iq, o.decimation)
    si = np.zeros(M, np.float32); sq = np.zeros(M, np.float32)
    L.orc_frontend(cu8, n_iq, C.byref(o), chain, si, sq)
    raw = np.zeros(M, np.float32)
    L.orc_discriminator(si, sq, M, int(o.accurate_atan), raw)
    fir = np.zeros(M, np.float32)
    L.orc_fir(raw, M, chain, fir)
    dphi = fir.copy()
    if o.remove_dc:
        L.orc_dcblock(dphi, M)
    bit = np.zeros(M, np.uint8); L.orc_slicer(dphi, M, bit)
    rssi = np.zeros(M, np.float32); L.orc_rssi(si, sq, M, rssi)
    clk = np.zeros(M, np.uint8); L.orc_clock(dphi, M, chain, clk)
    strobe = np.zeros(M, np.uint8); L.orc_time2_strobe(clk, M, strobe)
    return dict(M=M, si=si, sq=sq, dphi_raw=raw, fir=fir, dphi=dphi, bit=bit, rssi=rssi, clk=clk,
                strobe=strobe)


def events(st, chain, algo):
    L = lib()
    M = st["M"]
    cap = M // 2 + 64
    ev = np.zeros(cap, EVENT_DTYPE)
    if algo == 1:
        n = L.orc_time2_events(st["bit"], st["strobe"], st["rssi"], M, chain, ev.ctypes.data, cap)
    else:
        n = L.orc_runlength_events(st["bit"], st["rssi"], M, chain, ev.ctypes.data, cap)
    assert n <= cap
    return ev[:n]


def run_lines(cu8, o):
    L = lib()
    cu8 = np.ascontiguousarray(cu8, dtype=np.uint8)
    cap = 1 << 22
    buf = C.create_string_buffer(cap)
    nl = C.c_size_t(0)
    n = L.orc_run(cu8, len(cu8), C.byref(o), buf, cap, C.byref(nl))
    assert n < cap
    txt = buf.raw[:n].decode()
    return [l for l in txt.split("\n") if l]


def blank_ts(line):
    """Blank the TIMESTAMP column (4th, or 5th with the -v prefix)."""
    f = line.split(";")
    idx = 4 if f[0] in ("rla", "t2a") else 3
    f[idx] = "TS"
    return ";".join(f)


def ref_binary():
    p = os.path.join(REF_DIR, "rtl_wmbus")
    return p if os.path.exists(p) else None


def ref_lines(cu8_bytes, flags):
    """Lines from the UNMODIFIED reference binary (timestamps blanked)."""
    out = subprocess.run([ref_binary()] + flags.split(), input=bytes(cu8_bytes), capture_output=True,
                         check=True).stdout.decode()
    return [blank_ts(l) for l in out.split("\n") if l]


def ref_stage_dump(cu8_bytes, chain, o):
    exe = os.path.join(REF_DIR, "ref_stages")
    out = subprocess.run([exe, str(chain), str(o.decimation), str(int(o.accurate_atan)),
                          str(int(o.remove_dc)), str(int(o.simultaneous)), str(int(o.prefilter))],
                         input=bytes(cu8_bytes), capture_output=True, check=True).stdout
    return np.frombuffer(out, np.float32).reshape(-1, 6)
