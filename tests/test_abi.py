"""The C-ABI library must build, load on a CPU-only box, and export every symbol the header declares
(no compute call is made here - that needs a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry
    entry.build()
    from psa_b200 import _lib
    return _lib.load()


def _declared():
    text = (ROOT / "include" / "psa_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psa_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from psa_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/psa_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "ctypes signature table and header disagree"


def test_version_error_and_pitch_need_no_gpu(lib):
    assert lib.psa_version() >= 100
    assert isinstance(lib.psa_last_error(), bytes)
    assert lib.psa_pitch(1) == 64 and lib.psa_pitch(64) == 64 and lib.psa_pitch(65) == 128
    assert lib.psa_pitch(4096) == 4096


def test_bad_arguments_are_reported_not_crashed(lib):
    from psa_b200 import _lib
    # argument validation happens before any CUDA call, so this is safe without a device
    status = lib.psa_project(None, 0, 0, None, None, 0, 0, 0, None, 0, 0, None)
    assert status == _lib.ERR_BAD_ARG and b"psa_project" in lib.psa_last_error()
    with pytest.raises(ValueError):
        _lib.check(lib.psa_digitize(None, None, None, None, 1, 1, 1, 64, None, None, None))
    # strided row copy: empty extents are a no-op, a row wider than its pitch is refused
    assert lib.psa_copy_rows(None, 0, None, 0, 0, 0, None) == 0
    assert lib.psa_copy_rows(1, 8, 1, 16, 12, 4, None) == _lib.ERR_BAD_ARG and b"psa_copy_rows" in lib.psa_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from psa_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("PSA_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
    monkeypatch.delenv("PSA_B200_LIB")
    monkeypatch.setattr(_lib, "_lib", None)
    _lib.load()


def test_product_code_never_imports_the_oracle():
    for path in (ROOT / "psa_b200").rglob("*.py"):
        text = path.read_text()
        assert "oracle" not in re.sub(r'""".*?"""', "", text, flags=re.S).replace("# oracle", ""), path


def test_fft_planner_needs_no_gpu(lib):
    """Which frame counts take the mixed-radix core (no workspace) and which fall back to Bluestein is decided on
    the host: n = R * m with m = 4 * 2^a 3^b 5^c <= 4096 and R <= 256 is direct (R is a load-time split, so 28 =
    7 * 4 qualifies), anything else needs the float64 chirp scratch (16 bytes x 3 n_k x M, M = 2^s >= 2n-1)."""
    direct = [4, 8, 12, 16, 20, 28, 48, 60, 1000, 1200, 3000, 4096, 10000, 12288, 20000, 50000, 65536, 2 ** 19]
    for n in direct:
        assert lib.psa_fft_workspace_bytes(n, 7, 2) == 0, n
        assert lib.psa_fft_plan_bytes(n) >= 16 * n, n
    for n, m in ((1, 32), (2, 32), (7, 32), (250, 512), (3001, 8192), (8191, 16384), (4 * 2503, 32768), (10001, 32768)):
        assert lib.psa_fft_workspace_bytes(n, 7, 2) == 2 * 7 * 3 * m * 16, n
        assert lib.psa_fft_plan_bytes(n) >= 16 * (3 * m + n), n
    # 8192 / 16384 / 32768 frames take the four-step kernel: a ring of L2-resident float64 group slots
    # (16 columns x n_t x 16 bytes each, a few tens of MB whatever the k count) + tile counters
    for n in (8192, 16384, 32768):
        ws = lib.psa_fft_workspace_bytes(n, 1000, 1)
        assert 16 * n * 16 * 3 < ws < 64 << 20, (n, ws)
        assert lib.psa_fft_workspace_bytes(n, 2, 1) <= 16 * n * 16 + 4096, n       # one group: one slot
    for bad in (0, -3, 2 ** 19 + 1):
        assert lib.psa_fft_plan_bytes(bad) == -1


def test_balanced_digit_byte_trick():
    """The digitiser's carry-free digit extraction: the bytes of (X + 0x808080) ^ 0x808080 are the balanced
    base-256 digits of X for |X| <= 2^30 (csrc/ingest.cu: quad_words)."""
    import sys

    import numpy as np
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import intmodel as M
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.integers(-2 ** 30, 2 ** 30 + 1, 200000),
                        [2 ** 30, -2 ** 30, 0, 127, 128, -128, -129, 32767, 32768, -32768, -32769, 8388607, 8388608]])
    z = ((x + 0x00808080) & 0xFFFFFFFF) ^ 0x00808080
    digits = np.stack([(((z >> (8 * i)) & 0xFF) + 128) % 256 - 128 for i in range(4)]).astype(np.int8)
    np.testing.assert_array_equal(digits, M.balanced_digits(x.astype(np.int64)))
