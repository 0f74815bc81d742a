"""The C-ABI shared library: builds, loads, and exports exactly what include/gdeconv.h declares (CPU-only checks;
no compute call is made here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'gdeconv.h')).read()
    return sorted(set(re.findall(r'GD_API\s+[\w\s\*]+?\b(gd_\w+)\s*\(', text)))


def test_header_declares_and_library_exports_every_symbol():
    from gdeconv import _lib
    syms = _header_symbols()
    assert syms == sorted(_lib.SYMBOLS), 'include/gdeconv.h and gdeconv/_lib.py disagree'
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f'{s} is declared in include/gdeconv.h but not exported by libgdeconv.so'


def test_version_and_error_string_without_a_gpu():
    from gdeconv import _lib
    assert _lib.lib.gd_version() == 100
    assert _lib.lib.gd_workspace_bytes(0, 1, 64) > 64 * 2304 * 4
    assert _lib.lib.gd_workspace_bytes(7, 1, 64) == 0           # unknown arch is rejected, not guessed
    assert isinstance(_lib.lib.gd_last_error(), bytes)


def test_header_cites_the_reference_interfaces():
    text = open(os.path.join(ROOT, 'include', 'gdeconv.h')).read()
    for cite in ('models/unrolled_admm_gaussian.py:117-152', 'models/Unrolled_ADMM.py:177-215', 'models/ResUNet.py:26-42',
                 'models/Richard_Lucy.py:10-24', 'models/Wiener.py:10-20', 'models/Tikhonet.py:15-31',
                 'utils/utils_torch.py:46-50'):
        assert cite in text


def test_library_is_sm100a_native():
    """cuobjdump must show tcgen05 MMA, TMEM loads and bulk-async (TMA engine) copies in the shipped .so."""
    import shutil
    import subprocess
    from gdeconv import _lib
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass
    # tcgen05.mma, TMEM loads, bulk-async copies; cluster-multicast weight stages + multicast tcgen05.commit + cluster barrier
    for mnemonic in ('UTCHMMA', 'LDTM', 'UBLKCP', 'UBLKCP.S.G.MULTICAST', 'UTCBAR.MULTICAST', 'UCGABAR_ARV'):
        assert mnemonic in sass, mnemonic


def test_row_decode_division_is_exact_up_to_the_largest_admitted_chunk():
    """div_by_magic (csrc/gd_common.cuh) is exact only below 2^26 rows.  The library must (a) be exact on that
    whole range for every divisor the geometry uses (S and Wp of the four levels), and (b) refuse chunks that would exceed it."""
    import os
    from gdeconv import _lib, engine
    lib = _lib.lib
    for H in (48, 24, 12, 6):
        for d in ((H + 1) * (H + 1), H + 1):
            assert lib.gd_debug_divmagic(d, 0, 1 << 26) == -1, d
    cmax = lib.gd_max_chunk()
    assert 27000 < cmax < 28000
    for arch in (0, 1):
        assert lib.gd_workspace_bytes(arch, 1, cmax) > 0
        assert lib.gd_workspace_bytes(arch, 1, cmax + 1) == 0          # rejected, not mis-decoded
    old = os.environ.get('GDECONV_CHUNK')
    os.environ['GDECONV_CHUNK'] = '1000000'
    try:
        assert engine.max_chunk() == cmax
    finally:
        if old is None:
            del os.environ['GDECONV_CHUNK']
        else:
            os.environ['GDECONV_CHUNK'] = old


def test_fft_core_on_the_host(tmp_path):
    """csrc/fft_core.cuh is __host__ __device__: the phase functions the kernels run (and the in-register 48-point transform of
    k_wiener48) are compiled for the host and checked against the DFT definition in double precision (tests/cpu/test_fft_core.cpp)."""
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not found')
    exe = str(tmp_path / 'test_fft_core')
    r = subprocess.run([nvcc, '-O1', '-std=c++17', '-Wno-deprecated-gpu-targets', '-I', os.path.join(ROOT, 'galaxy-deconv_b200', 'csrc'),
                        os.path.join(ROOT, 'tests', 'cpu', 'test_fft_core.cpp'), '-o', exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith('OK'), r.stdout[-2000:]
