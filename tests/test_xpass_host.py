"""The fused x pass (csrc/trvb_xpass.cuh: forward FFT along x, low-|k| modes, shot-noise
spectrum, inverse FFT along x) emulated on the CPU: the header is plain C++ with the thread
index as an argument, so its stage sequence, digit reversal, twiddles and low-|k| indexing
are checked here against numpy for every supported length without a GPU.  The arithmetic it
replaces: S/field.cpp:1496-1655 (x pass of the forward transform), 3273-3298 (spectrum),
3318-3345 (x pass of the inverse transform)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = tmp_path_factory.mktemp("xpass") / "xpass_host"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", f"-I{ROOT / 'triumvirate_b200' / 'csrc'}",
                    str(ROOT / "tests" / "native" / "xpass_host.cpp"), "-o", str(exe)], check=True)
    return exe


@pytest.mark.parametrize("n0,n1,n2,sub", [
    (32, 8, 8, (6, 6, 6)), (64, 8, 6, (10, 4, 4)), (128, 4, 8, (20, 4, 6)), (256, 8, 4, (0, 0, 0)),
    (512, 4, 6, (144, 4, 4)), (1024, 4, 4, (626, 2, 4)), (2048, 2, 4, (0, 0, 0)),
    (64, 3, 5, (8, 3, 3)),   # ragged: n1 nh not a multiple of the columns per tile
])
@pytest.mark.parametrize("ck4", [False, True])
@pytest.mark.parametrize("load", ["registers", "cp.async"])
def test_fused_x_pass_matches_numpy(harness, tmp_path, monkeypatch, n0, n1, n2, sub, ck4, load):
    if load == "cp.async":
        monkeypatch.setenv("XP_ASYNC", "1")   # tile copied first, stage 0 run like a middle stage
    if ck4:
        if n0 != 512:
            pytest.skip("four-column tiles are an option of the 512 kernel only")
        monkeypatch.setenv("XP_CK4", "1")   # the 4-column tile (bank swizzle with radix-8 junction)
    gen = np.random.default_rng(n0 + n1)
    nh = n2 // 2 + 1
    T = gen.normal(size=(n0, n1, nh)) + 1j * gen.normal(size=(n0, n1, nh))
    ral = [gen.uniform(1., 2., size=n) for n in (n0, n1, nh)]
    add_a, add_b, S, inv_vol = -3.5, 1.25, 0.75, 1. / 7.
    s0, s1, s2 = sub
    T.astype(np.complex128).tofile(tmp_path / "in.bin")
    np.concatenate(ral).tofile(tmp_path / "ral.bin")
    subprocess.run([str(harness), str(n0), str(n1), str(nh), str(s0), str(s1), str(s2),
                    repr(add_a), repr(add_b), repr(S), repr(inv_vol), str(tmp_path / "in.bin"),
                    str(tmp_path / "ral.bin"), str(tmp_path / "out.bin"), str(tmp_path / "lowk.bin")],
                   check=True)
    out = np.fromfile(tmp_path / "out.bin", dtype=np.complex128).reshape(n0, n1, nh)
    U = np.fft.fft(T, axis=0)
    a, b = U.copy(), U.copy()
    a[0, 0, 0] += add_a
    b[0, 0, 0] += add_b
    rc1 = ral[0][:, None, None] * ral[1][None, :, None] * ral[2][None, None, :]
    W = (a * np.conj(b) * rc1 - S) * inv_vol
    want = np.fft.ifft(W, axis=0) * n0
    assert np.max(np.abs(out - want)) < 1.e-12 * np.max(np.abs(want))
    if s0:
        sh = s2 // 2 + 1
        lowk = np.fromfile(tmp_path / "lowk.bin", dtype=np.complex128).reshape(s0, s1, sh)
        ref = np.zeros_like(lowk)
        for is_ in range(s0):
            mi = is_ if is_ < s0 // 2 else is_ - s0
            if 2 * abs(mi) >= s0:
                continue
            for js in range(s1):
                mj = js if js < s1 // 2 else js - s1
                if 2 * abs(mj) >= s1:
                    continue
                for ks in range(sh):
                    if 2 * ks >= s2:
                        continue
                    ref[is_, js, ks] = a[mi % n0, mj % n1, ks]
        assert np.max(np.abs(lowk - ref)) < 1.e-12 * np.max(np.abs(a))


@pytest.fixture(scope="module")
def zharness(tmp_path_factory):
    exe = tmp_path_factory.mktemp("zpass") / "zpass_host"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{ROOT / 'triumvirate_b200' / 'csrc'}",
                    str(ROOT / "tests" / "native" / "zpass_host.cpp"), "-o", str(exe)], check=True)
    return exe


ZPASS_LENGTHS = [64, 72, 96, 108, 128, 144, 160, 180, 192, 216, 240, 256, 270, 288, 320, 360, 384,
                 432, 480, 512, 540, 576, 600, 640, 720]


def test_zpass_length_list_matches_the_device_layer():
    """The lengths emulated here are the ones csrc/trvb_fourier.cu dispatches."""
    import re
    src = (ROOT / "triumvirate_b200" / "csrc" / "trvb_fourier.cu").read_text()
    block = src[src.index("#define TRVB_ZPASS_LENGTHS(X)"):]
    block = block[:block.index("\n\n")]
    assert sorted(int(v) for v in re.findall(r"X\((\d+)\)", block)) == ZPASS_LENGTHS


@pytest.mark.parametrize("n", ZPASS_LENGTHS)
def test_pruned_c2r_z_pass_matches_numpy(zharness, tmp_path, n):
    """csrc/trvb_zpass.cuh (last pass of the pruned shell transform, S/field.cpp:1792-1906):
    two real lines packed into one complex line, digit-reversed load, mixed-radix
    decimation-in-time stages, natural-order store -- against numpy.fft.irfft on zero-padded
    lines, with a ragged number of lines."""
    gen = np.random.default_rng(n)
    n1 = 37 if n % 16 == 0 else 22
    k2 = n // 4 if n != 540 else 135
    B = gen.normal(size=(k2, n1)) + 1j * gen.normal(size=(k2, n1))
    B.astype(np.complex128).tofile(tmp_path / "in.bin")
    subprocess.run([str(zharness), str(n), str(n1), str(k2), str(tmp_path / "in.bin"),
                    str(tmp_path / "out.bin")], check=True)
    out = np.fromfile(tmp_path / "out.bin").reshape(n1, n)
    F = np.zeros((n1, n // 2 + 1), dtype=complex)
    F[:, :k2] = B.T
    F[:, 0] = F[:, 0].real     # a c2r transform ignores Im F[0]
    want = np.fft.irfft(F, n=n, axis=1) * n
    assert np.max(np.abs(out - want)) < 1.e-13 * np.max(np.abs(want))


@pytest.fixture(scope="module")
def yharness(tmp_path_factory):
    exe = tmp_path_factory.mktemp("ypass") / "ypass_host"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{ROOT / 'triumvirate_b200' / 'csrc'}",
                    str(ROOT / "tests" / "native" / "ypass_host.cpp"), "-o", str(exe)], check=True)
    return exe


@pytest.mark.parametrize("n", ZPASS_LENGTHS)
def test_pruned_c2c_y_pass_matches_numpy(yharness, tmp_path, n):
    """csrc/trvb_zpass.cuh, y pass of the pruned shell transform: K1 = 2 mc + 1 modes per line
    loaded along x to digit-reversed slots, decimation-in-time stages, natural-order store
    for a window of planes [x0, x0 + nx) -- against numpy.fft.ifft on zero-padded lines.  One
    length runs the unbounded case (every mode below the Nyquist frequency: the G field)."""
    gen = np.random.default_rng(n + 1)
    n0, x0, nx = 29, 3, 10
    mc1 = (n - 1) // 2 if n == 72 else n // 4 - 1
    k1 = 2 * mc1 + 1
    A = gen.normal(size=(k1, n0)) + 1j * gen.normal(size=(k1, n0))
    A.astype(np.complex128).tofile(tmp_path / "in.bin")
    subprocess.run([str(yharness), str(n), str(n0), str(mc1), str(x0), str(nx),
                    str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], check=True)
    out = np.fromfile(tmp_path / "out.bin", dtype=np.complex128).reshape(nx, n)
    F = np.zeros((nx, n), dtype=complex)
    for b in range(k1):
        F[:, (b - mc1) % n] = A[b, x0:x0 + nx]
    want = np.fft.ifft(F, axis=1) * n
    assert np.max(np.abs(out - want)) < 1.e-13 * np.max(np.abs(want))


def test_library_reports_the_same_z_pass_lengths():
    """trvb_shell_zpass_supported (a host-only query of libtrvb.so) agrees with the list the
    emulation covers, and nothing else in 2 .. 800 is claimed."""
    from triumvirate_b200 import _lib
    lib = _lib.trvb()
    got = [n for n in range(2, 801) if lib.trvb_shell_zpass_supported(n) == 1]
    assert got == ZPASS_LENGTHS
