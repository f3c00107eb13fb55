"""The C-ABI shared library loads without a GPU, exports every symbol include/metabuli_b200.h declares, and
fails loudly (MBL_E_NO_DEVICE) instead of falling back when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "metabuli_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mbl_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_exported():
    from metabuli_b200 import _ffi
    lib = _ffi.load_library()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(_ffi.EXPORTS) == set(names)


def test_struct_layouts():
    from metabuli_b200 import _ffi
    assert C.sizeof(_ffi.ReadResult) == 28
    assert _ffi.MATCH_DTYPE.itemsize == 24
    assert C.sizeof(_ffi.Stats) == 4 * 8 + 3 * 8 + 4 * 4 + 2 * 4 + 8 + 2 * 4 + 2 * 4
    assert C.sizeof(_ffi.Config) == 16 * 4
    assert C.sizeof(_ffi.Shard) == 56


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from metabuli_b200 import _ffi
    lib = _ffi.load_library()
    cfg = _ffi.Config(kmer_format=2, seq_mode=1, tie_ratio=0.95, min_cons_cnt=4, min_cons_cnt_euk=9, match_per_kmer=4)
    ctx = C.c_void_p()
    assert lib.mbl_create(C.byref(cfg), C.byref(ctx)) == _ffi.MBL_E_NO_DEVICE
    assert not ctx.value


def test_product_does_not_touch_oracle():
    """Nothing under metabuli_b200/ may reference the oracle (rule ③)."""
    for base, _, files in os.walk(os.path.join(ROOT, "metabuli_b200")):
        if "_lib" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "mbl_oracle" not in txt and "import oracle" not in txt and "/oracle/" not in txt, os.path.join(base, f)


def test_cli_rejects_flags_that_would_change_the_output():
    """`--reduced-aa 1` changes the reference's classifications and is not implemented on the B200 path: the C++ host must die
    with a message, not drop it (argument parsing needs no GPU).  (`--mask 1` is implemented since round 2:
    tests/test_host_mask.py, test_gpu_synth.py::test_masked_queries; `--taxonomy-path` is accepted: the reference ignores it whenever
    <dbdir>/taxonomyDB exists, which this host requires — test_cli_cpu.py::test_taxonomy_path_is_ignored_like_in_the_reference.)"""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "metabuli_b200", "_lib", "metabuli-b200")
    for flags in (["--reduced-aa", "1"], ["--em"], ["--em", "1"], ["--no-such-flag", "1"]):
        r = subprocess.run([exe, "classify", "--seq-mode", "1", *flags, "a.fna", "db", "out", "job"], capture_output=True, text=True)
        assert r.returncode != 0 and "Error" in (r.stdout + r.stderr), flags
    # harmless spellings are accepted up to the input checks
    r = subprocess.run([exe, "classify", "--seq-mode", "1", "--mask", "1", "--mask-prob", "0.8", "--em", "0", "--syncmer", "0", "--smer-len", "5", "--kmer-format", "1", "--print-log", "0", "--taxonomy-path", "x", "--max-ram", "8", "--hamming-margin", "1", "a.fna", "db", "out", "job"],
                       capture_output=True, text=True)
    assert "not supported" not in (r.stdout + r.stderr)
