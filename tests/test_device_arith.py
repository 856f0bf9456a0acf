"""The two places where the device code leaves the reference's literal operation sequence are proven in comments;
this checks the proofs numerically on the CPU build of the same source (`not gpu`):
 * wmb_atan2f_t<true> (range escapes of fdlibm's atan2f dropped, zero cases by select) against the oracle's full
   restatement on the discriminator's whole input domain corners: multiples of 1/len^2 with numerators below 2^23;
 * wmb_div_small (multiply-high instead of a divide) against C truncating division."""
import ctypes as C

import numpy as np


def test_bounded_atan2f_equals_full_on_the_discriminator_domain(hostsim_lib, orc_mod):
    L = orc_mod.lib()
    hb = hostsim_lib.hostsim_atan2f_bounded
    hg = hostsim_lib.hostsim_atan2f_general
    for f in (hb, hg):
        f.argtypes = [C.c_float, C.c_float]; f.restype = C.c_float
    rng = np.random.default_rng(23)
    edge = np.array([0, 1, -1, 2, -2, 3, 64, -64, 255, 4095, -4096, (1 << 23) - 1, -(1 << 23) + 1, 2 * 2032 * 2032, -2 * 2032 * 2032,
                     2 * 1016 * 1016, 7, -7], np.int64)
    ys = np.concatenate([np.repeat(edge, len(edge)), rng.integers(-(1 << 23) + 1, 1 << 23, 40000),
                         rng.integers(-3000, 3000, 40000)])
    xs = np.concatenate([np.tile(edge, len(edge)), rng.integers(-(1 << 23) + 1, 1 << 23, 40000),
                         rng.integers(-3000, 3000, 40000)])
    bad = 0
    for scale in (1.0 / 64.0, 1.0 / 256.0):
        for y, x in zip(ys, xs):
            yf, xf = np.float32(y * scale), np.float32(x * scale)
            want = np.float32(L.orc_atan2f(yf, xf)).view(np.uint32)
            got = np.float32(hb(yf, xf)).view(np.uint32)
            gen = np.float32(hg(yf, xf)).view(np.uint32)
            bad += int(got != want) + int(gen != want)
    assert bad == 0


def test_div_small_is_c_division(hostsim_lib):
    f = hostsim_lib.hostsim_div_small
    f.argtypes = [C.c_int, C.c_int]; f.restype = C.c_int
    rng = np.random.default_rng(5)
    xs = np.concatenate([np.arange(-5000, 5000), rng.integers(-(1 << 27), 1 << 27, 20000),
                         np.array([(1 << 27) - 1, -(1 << 27) + 1, (1 << 26), -(1 << 26)])])
    for n in range(1, 9):
        for x in xs:
            x = int(x)
            want = abs(x) // n * (1 if x >= 0 else -1)             # truncation toward zero
            assert f(x, n) == want, (x, n)
    g = hostsim_lib.hostsim_div_pow2
    g.argtypes = [C.c_int, C.c_int]; g.restype = C.c_int
    for s in (4, 5):
        for x in xs[:12000]:
            x = int(x)
            assert g(x, s) == abs(x) // (1 << s) * (1 if x >= 0 else -1), (x, s)
