"""Product framers (rtl-wmbus_b200/csrc/wmb_framer.c, frame-at-once) vs the oracle's per-bit state
machines (restating t1_c1_packet_decoder.h / s1_packet_decoder.h) on random and crafted bit lists."""
import ctypes as C
import importlib

import numpy as np
import pytest


class Decoded(C.Structure):
    _fields_ = [("status", C.c_int), ("consumed", C.c_uint32), ("end_sample", C.c_uint64), ("mode", C.c_char * 3),
                ("crc_ok", C.c_uint8), ("ok_3of6", C.c_uint8), ("packet_rssi", C.c_uint32),
                ("current_rssi", C.c_uint32), ("serial", C.c_uint32), ("len", C.c_uint32),
                ("datagram", C.c_uint8 * 292)]


def product_decode(lib, pkg, chain, bits, rssi):
    n = len(bits)
    words = np.zeros(n, np.uint32)
    words[:] = (np.arange(n, dtype=np.uint32) << 9) | (rssi.astype(np.uint32) << 1) | bits.astype(np.uint32)
    f = pkg.WmbFrame()
    f.sync_sample = 1000; f.ordinal = 5; f.chain = chain; f.algo = 0; f.nbits = n
    f.bits = words.ctypes.data_as(C.POINTER(C.c_uint32))
    d = Decoded()
    lib.wmb_frame_decode.argtypes = [C.POINTER(pkg.WmbFrame), C.POINTER(Decoded)]
    lib.wmb_frame_decode(C.byref(f), C.byref(d))
    line = None
    if d.status == 1:
        buf = C.create_string_buffer(2048)
        lib.wmb_format_line.argtypes = [C.POINTER(Decoded), C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
        lib.wmb_format_line.restype = C.c_size_t
        k = lib.wmb_format_line(C.byref(d), b"rla;", b"TS", buf, 2048)
        line = buf.raw[:k].decode().rstrip("\n")
    return d.status, d.consumed, line, d.end_sample


def oracle_decode(orc_mod, chain, bits, rssi):
    L = orc_mod.lib()
    out = C.create_string_buffer(4096)
    got = C.c_int(0)
    fn = L.orc_frame_t1c1 if chain == 0 else L.orc_frame_s1
    consumed = fn(np.ascontiguousarray(bits, np.uint8), np.ascontiguousarray(rssi, np.uint8), len(bits), b"rla;",
                  out, 4096, C.byref(got))
    line = out.value.decode().rstrip("\n") if got.value else None
    return consumed, line


def frames_for_tests(synth):
    e = synth.Emitter("T1", 0x12345678, l_field=0x1E, seed=5)
    p = e.payload(3)
    out = []
    a = synth.frame_a(p)
    t1 = synth.chips_t1(a, preamble_pairs=0, post_pairs=8)[len(synth.SYNC_T1C1) - 1:]
    c1a = synth.chips_c1(a, False, preamble_pairs=0, post_pairs=8)[len(synth.SYNC_T1C1) - 1:]
    c1b = synth.chips_c1(synth.frame_b(p), True, preamble_pairs=0, post_pairs=8)[len(synth.SYNC_T1C1) - 1:]
    s1 = synth.chips_s1(a, preamble_pairs=0, post_pairs=8)[len(synth.SYNC_S1) - 1:]
    big = synth.Emitter("C1B", 0x00112233, l_field=0xF0, seed=6).payload(1)
    c1b_big = synth.chips_c1(synth.frame_b(big), True, preamble_pairs=0, post_pairs=8)[len(synth.SYNC_T1C1) - 1:]
    t1_max = synth.chips_t1(synth.frame_a(synth.Emitter("T1", 1, l_field=0xFF, seed=7).payload(0)), 0, 8)[len(synth.SYNC_T1C1) - 1:]
    return [(0, t1), (0, c1a), (0, c1b), (1, s1), (0, c1b_big), (0, t1_max)]


def test_framer_on_clean_and_corrupted_frames(hostsim_lib, pkg, orc_mod):
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    rng = np.random.default_rng(3)
    n_lines = 0
    for chain, chips in frames_for_tests(synth):
        chips = chips.astype(np.uint8)
        for trial in range(60):
            bits = chips.copy()
            rssi = rng.integers(20, 200, len(bits)).astype(np.uint8)
            if trial % 3 == 1:                      # flip a few chips
                for i in rng.integers(1, len(bits), rng.integers(1, 4)):
                    bits[i] ^= 1
            if trial % 3 == 2:                      # rssi drop-outs -> PACKET_CAPTURE_THRESHOLD abort
                rssi[rng.integers(0, len(bits))] = rng.integers(0, 5)
            if trial % 10 == 9:                     # truncated list
                cut = rng.integers(1, len(bits))
                bits, rssi = bits[:cut], rssi[:cut]
            st, consumed, line, end = product_decode(hostsim_lib, pkg, chain, bits, rssi)
            oc, oline = oracle_decode(orc_mod, chain, bits, rssi)
            if st == 2:                             # NEED_MORE: the oracle ran through the whole list
                assert oline is None and oc == len(bits)
            else:
                assert consumed == oc, (chain, trial, consumed, oc)
                assert line == oline, (chain, trial)
                if line:
                    n_lines += 1
                    assert end == 1000 + consumed - 1
    assert n_lines > 100


def test_framer_on_random_bits(hostsim_lib, pkg, orc_mod):
    rng = np.random.default_rng(11)
    for trial in range(3000):
        chain = trial & 1
        n = int(rng.integers(1, 400))
        bits = rng.integers(0, 2, n).astype(np.uint8)
        if trial % 4 == 0 and chain == 0 and n > 30:     # force the C1 mode word path
            word = "010101001100" if trial % 8 == 0 else "010101000011"
            bits[1:13] = np.frombuffer(word.encode(), np.uint8) - 48
            bits[13:17] = [1, 1, 0, 1]
            bits[17:25] = [0, 0, 0, 0, 0, 0, int(rng.integers(0, 2)), int(rng.integers(0, 2))]
        rssi = rng.integers(3, 60, n).astype(np.uint8)
        st, consumed, line, _ = product_decode(hostsim_lib, pkg, chain, bits, rssi)
        oc, oline = oracle_decode(orc_mod, chain, bits, rssi)
        if st == 2:
            assert oline is None and oc == n
        else:
            assert (consumed, line) == (oc, oline), (trial, chain)


# ---- device framer (kernel K4) vs its host twin, candidate by candidate ----------------------------

def framer_cases(synth, n_random=1500):
    """(chain, bits, rssi) of crafted, corrupted, truncated and random candidates."""
    rng = np.random.default_rng(17)
    cases = []
    for chain, chips in frames_for_tests(synth):
        chips = chips.astype(np.uint8)
        for trial in range(40):
            bits = chips.copy()
            rssi = rng.integers(20, 200, len(bits)).astype(np.uint8)
            if trial % 4 == 1:
                for i in rng.integers(1, len(bits), rng.integers(1, 4)):
                    bits[i] ^= 1
            if trial % 4 == 2:
                rssi[rng.integers(0, len(bits))] = rng.integers(0, 5)
            if trial % 4 == 3:                      # low rssi exactly on / next to the telegram's last bit
                rssi[len(bits) - 17 + int(rng.integers(-1, 2))] = 1
            if trial % 10 == 9:
                cut = rng.integers(1, len(bits))
                bits, rssi = bits[:cut], rssi[:cut]
            cases.append((chain, bits, rssi))
    for trial in range(n_random):
        chain = trial & 1
        n = int(rng.integers(1, 400))
        bits = rng.integers(0, 2, n).astype(np.uint8)
        if trial % 4 == 0 and chain == 0 and n > 30:
            word = "010101001100" if trial % 8 == 0 else "010101000011"
            bits[1:13] = np.frombuffer(word.encode(), np.uint8) - 48
            bits[13:17] = [1, 1, 0, 1]
            bits[17:25] = [0, 0, 0, 0, 0, 0, int(rng.integers(0, 2)), int(rng.integers(0, 2))]
        if trial % 4 == 1 and chain == 1:                # valid Manchester prefix of random length
            k = int(rng.integers(1, n // 2 + 1))
            v = rng.integers(0, 2, k).astype(np.uint8)
            bits[1:1 + 2 * k:2][:len(bits[1:1 + 2 * k:2])] = (1 - v)[:len(bits[1:1 + 2 * k:2])]
            bits[2:2 + 2 * k:2][:len(bits[2:2 + 2 * k:2])] = v[:len(bits[2:2 + 2 * k:2])]
        rssi = rng.integers(3, 60, n).astype(np.uint8)
        cases.append((chain, bits, rssi))
    return cases


def check_device_framer(lib, pkg, synth, orc_mod=None):
    """K4 against its host twin, field by field, and -- third column -- against the oracle's per-bit state machines
    (consumed bits and the formatted line) on the same candidates."""
    cases = framer_cases(synth)
    n = len(cases)
    frames = (pkg.WmbFrame * n)()
    keep = []
    for i, (chain, bits, rssi) in enumerate(cases):
        k = len(bits)
        words = (np.arange(k, dtype=np.uint32) * 3 << 9) | (rssi.astype(np.uint32) << 1) | bits.astype(np.uint32)
        words = np.ascontiguousarray(words, np.uint32)
        keep.append(words)
        f = frames[i]
        f.sync_sample = 1000 + i; f.ordinal = i; f.chain = chain; f.algo = i & 1; f.nbits = k
        f.bits = words.ctypes.data_as(C.POINTER(C.c_uint32))
    host = (Decoded * n)()
    dev = (Decoded * n)()
    lib.wmb_frame_decode.argtypes = [C.POINTER(pkg.WmbFrame), C.POINTER(Decoded)]
    for i in range(n):
        lib.wmb_frame_decode(C.byref(frames[i]), C.byref(host[i]))
    with pkg.WmbusB200("-v", lib=lib) as ctx:
        lib.wmb_frame_decode_device.argtypes = [C.c_void_p, C.POINTER(pkg.WmbFrame), C.c_size_t, C.POINTER(Decoded)]
        rc = lib.wmb_frame_decode_device(ctx._ctx, frames, n, dev)
        assert rc == 0, lib.wmb_last_error()
    seen = {0: 0, 1: 0, 2: 0}
    for i in range(n):
        h, d = host[i], dev[i]
        assert (h.status, h.consumed, h.end_sample) == (d.status, d.consumed, d.end_sample), (i, cases[i][0])
        seen[h.status] += 1
        if h.status == 1:
            assert (h.mode, h.crc_ok, h.ok_3of6, h.packet_rssi, h.current_rssi, h.serial, h.len) == \
                   (d.mode, d.crc_ok, d.ok_3of6, d.packet_rssi, d.current_rssi, d.serial, d.len), i
            assert bytes(h.datagram[:h.len]) == bytes(d.datagram[:d.len]), i
    assert seen[0] > 500 and seen[1] > 100 and seen[2] > 10, seen
    if orc_mod is None:
        return
    lib.wmb_format_line.argtypes = [C.POINTER(Decoded), C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
    lib.wmb_format_line.restype = C.c_size_t
    buf = C.create_string_buffer(2048)
    n_lines = 0
    for i, (chain, bits, rssi) in enumerate(cases):
        d = dev[i]
        oc, oline = oracle_decode(orc_mod, chain, bits, rssi)
        if d.status == 2:                               # NEED_MORE: the oracle ran through the whole list
            assert oline is None and oc == len(bits), i
            continue
        assert d.consumed == oc, (i, chain, d.consumed, oc)
        line = None
        if d.status == 1:
            k = lib.wmb_format_line(C.byref(d), b"rla;", b"TS", buf, 2048)
            line = buf.raw[:k].decode().rstrip("\n")
            n_lines += 1
        assert line == oline, (i, chain)
    assert n_lines > 100


def test_device_framer_matches_host_twin_and_oracle_hostsim(hostsim_lib, pkg, orc_mod):
    check_device_framer(hostsim_lib, pkg, importlib.import_module("rtl-wmbus_b200.synth"), orc_mod)


@pytest.mark.gpu
def test_device_framer_matches_host_twin_and_oracle_gpu(gpu_lib, pkg, orc_mod):
    check_device_framer(gpu_lib, pkg, importlib.import_module("rtl-wmbus_b200.synth"), orc_mod)


def test_time_string_format(hostsim_lib):
    import re
    buf = C.create_string_buffer(64)
    hostsim_lib.wmb_make_time_string.argtypes = [C.c_char_p, C.c_size_t]
    hostsim_lib.wmb_make_time_string.restype = None
    for _ in range(3):
        hostsim_lib.wmb_make_time_string(buf, 64)
        assert re.fullmatch(r"\d{4}-\d\d-\d\d \d\d:\d\d:\d\d\.\d{6}", buf.value.decode()), buf.value
    hostsim_lib.wmb_make_time_string(buf, 8)          # too small: empty string, no overflow
    assert buf.value == b""
