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


def _arith_operands(n, seed):
    """integer-valued operands as the discriminator meets them: small box sums, large ones, zeros, edges"""
    rng = np.random.default_rng(seed)
    edge = np.array([0, 1, -1, 2, -2, 3, 64, -64, 255, 4095, -4096, (1 << 23) - 1, -(1 << 23) + 1, 2 * 2032 * 2032,
                     -2 * 2032 * 2032, 2 * 1016 * 1016, 7, -7, 1 << 21, -(1 << 21)], np.int64)
    parts_y = [np.repeat(edge, len(edge)), rng.integers(-(1 << 23) + 1, 1 << 23, n), rng.integers(-3000, 3000, n),
               rng.integers(-40, 40, n), rng.integers(-(1 << 21), 1 << 21, n)]
    parts_x = [np.tile(edge, len(edge)), rng.integers(-(1 << 23) + 1, 1 << 23, n), rng.integers(-3000, 3000, n),
               rng.integers(-40, 40, n), rng.integers(-40, 40, n)]
    return np.concatenate(parts_y), np.concatenate(parts_x)


def check_device_arith(pkg, lib, orc_mod, n):
    """wmb_debug_arith: the device's atan2f (bounded and general), IEEE division and sqrt without the slow paths, and
    the discriminator, operand by operand against the oracle / numpy's correctly rounded float32 operations."""
    import orc
    L = orc_mod.lib()
    ys, xs = _arith_operands(n, 31)
    atan = np.vectorize(lambda a, b: L.orc_atan2f(float(a), float(b)), otypes=[np.float32])
    with pkg.WmbusB200("", lib=lib) as ctx:
        for scale in (1.0, 1.0 / 64.0, 1.0 / 256.0):
            yf, xf = (ys * scale).astype(np.float32), (xs * scale).astype(np.float32)
            want = atan(yf, xf)
            for mode in (0, 1):
                got = ctx.debug_arith(mode, yf, xf)
                bad = np.nonzero(got.view(np.uint32) != want.view(np.uint32))[0]
                assert len(bad) == 0, (mode, scale, len(bad), yf[bad[:4]], xf[bad[:4]], got[bad[:4]], want[bad[:4]])
            nz = xf != 0
            got = ctx.debug_arith(2, yf[nz], xf[nz])
            with np.errstate(all="ignore"):
                wantd = (yf[nz] / xf[nz]).astype(np.float32)
            bad = np.nonzero(got.view(np.uint32) != wantd.view(np.uint32))[0]
            assert len(bad) == 0, ("div", scale, len(bad), yf[nz][bad[:4]], xf[nz][bad[:4]])
        # the argument reduction's second division: (2t-1)/(2+t), (t-1)/(t+1), (t-1.5)/(1+1.5t), -1/t
        t = np.abs(np.random.default_rng(3).standard_cauchy(n)).astype(np.float32) + np.float32(0.4375)
        for num, den in ((2 * t - 1, 2 + t), (t - 1, t + 1), (t - np.float32(1.5), 1 + np.float32(1.5) * t), (np.full_like(t, -1), t)):
            num, den = num.astype(np.float32), den.astype(np.float32)
            got = ctx.debug_arith(2, num, den)
            wantd = (num / den).astype(np.float32)
            assert np.array_equal(got.view(np.uint32), wantd.view(np.uint32))
        # sqrt of i^2 + q^2, integers below 2^23 (and zero)
        a = np.concatenate([np.arange(0, 70000), np.random.default_rng(4).integers(0, 1 << 23, n)]).astype(np.float32)
        got = ctx.debug_arith(3, a, a)
        assert np.array_equal(got.view(np.uint32), np.sqrt(a).astype(np.float32).view(np.uint32))
        # the discriminator on consecutive (I, Q) pairs of box sums, scaled (len 8 / 16) and unscaled: identical
        si = np.random.default_rng(6).integers(-1016, 1017, n).astype(np.float32)
        sq = np.random.default_rng(7).integers(-1016, 1017, n).astype(np.float32)
        si[::97] = 0; sq[::89] = 0; si[5::1000] = 0; sq[5::1000] = 0            # zeros, also both at once
        ref = ctx.debug_arith(4, si / np.float32(8), sq / np.float32(8))
        raw = np.zeros(n, np.float32)
        L.orc_discriminator(np.ascontiguousarray(si / np.float32(8)), np.ascontiguousarray(sq / np.float32(8)), n, 1, raw)
        assert np.array_equal(ref[1:].view(np.uint32), raw[1:].view(np.uint32))
        for s in (np.float32(1), np.float32(1 / 16)):
            assert np.array_equal(ctx.debug_arith(4, si * s, sq * s)[1:].view(np.uint32), ref[1:].view(np.uint32))


def test_device_arith_hooks_cpu_build(pkg, hostsim_lib, orc_mod):
    check_device_arith(pkg, hostsim_lib, orc_mod, 20000)
