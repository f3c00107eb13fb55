"""`--mask 1`: the host-side tantan masking (mbl_mask_reads, csrc/host/tantan_mask.hpp) against per-letter goldens written by the
reference's own NucleotideMatrix / ProbabilityMatrix / tantan objects (oracle/ref_mask_main.cpp via tests/golden/gen_synth_golden.py).
Host work only: runs without a GPU."""
import ctypes as C
import gzip
import os

import numpy as np
import pytest

import synth_cases


def _mask(bases, offsets, prob, threads=2):
    from metabuli_b200 import _ffi
    lib = _ffi.load_library()
    b = bases.copy()
    o = np.ascontiguousarray(offsets, dtype=np.uint64)
    rc = lib.mbl_mask_reads(b.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), o.size - 1, C.c_float(prob), threads)
    assert rc == 0
    return b


def _lines(bases, offsets):
    return b"".join(bytes(bases[int(offsets[i]):int(offsets[i + 1])]) + b"\n" for i in range(offsets.size - 1))


def test_masked_letters_of_the_synthetic_case(golden_dir):
    sdb, reads, _ = synth_cases.build("mask_se")
    got = _mask(reads[0], reads[1], synth_cases.mask_flags("mask_se")[1], threads=3)
    want = gzip.open(os.path.join(golden_dir, "synth", "mask_se.masked.gz"), "rb").read()
    assert _lines(got, reads[1]) == want
    n_masked = int((got == ord("N")).sum() - (reads[0] == ord("N")).sum())
    assert n_masked > 50000                       # the case is not vacuous
    # unmasked letters are the caller's, case and IUPAC codes included
    keep = got != ord("N")
    assert np.array_equal(got[keep], reads[0][keep])


def test_odd_inputs_and_thresholds(golden_dir):
    b, o = synth_cases.mask_misc_reads()
    want = gzip.open(os.path.join(golden_dir, "synth", "mask_misc.masked.gz"), "rb").read()
    got = b"".join(_lines(_mask(b, o, p), o) for p in synth_cases.MASK_MISC_PROBS)
    assert got == want


@pytest.mark.parametrize("threads", [1, 5, 0])
def test_thread_count_does_not_change_the_mask(threads):
    # 0 = all cores; batches under 256 reads stay on one thread, so take a larger one
    sdb, reads, _ = synth_cases.build("mask_se")
    assert np.array_equal(_mask(reads[0], reads[1], 0.9, threads), _mask(reads[0], reads[1], 0.9, 1))


def test_empty_batch_and_bad_arguments():
    from metabuli_b200 import _ffi
    lib = _ffi.load_library()
    off = np.zeros(1, dtype=np.uint64)
    assert lib.mbl_mask_reads(None, off.ctypes.data_as(C.c_void_p), 0, C.c_float(0.9), 1) == 0
    assert lib.mbl_mask_reads(None, None, 0, C.c_float(0.9), 1) == _ffi.MBL_E_BAD_ARG


def test_device_kernel_source_under_cpu_emulation(golden_dir):
    """The source of K0 (csrc/k0_mask.cu) compiled for the CPU by tests/host/k0_emulation.cpp — one block of four warps as 128
    lock-step host threads — against the reference's per-letter golden: odd reads, tandem repeats around the 50-offset limit and
    reads of 1000 / 1001 / 4097 letters (the code window moves in both sweeps).  The GPU suite runs the same inputs on the device."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "host", "k0_emulation.cpp")
    out = os.path.join(root, "tests", "host", "_build", "k0_emulation")
    deps = [src, os.path.join(root, "metabuli_b200", "csrc", "k0_mask.cu"), os.path.join(root, "metabuli_b200", "csrc", "host", "tantan_model.hpp")]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(out) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off", "-mfma", "-pthread", src, "-o", out])
    b, o = synth_cases.mask_misc_reads()
    n_reads = 33                                     # up to and including the 4097-letter read; the 20 kb one is left to the GPU
    assert int(o[n_reads] - o[n_reads - 1]) == 4097
    lines = _lines(b, o[: n_reads + 1])
    want_all = gzip.open(os.path.join(golden_dir, "synth", "mask_misc.masked.gz"), "rb").read()
    n_line = int(o[-1]) + o.size - 1
    k = synth_cases.MASK_MISC_PROBS.index(0.9)
    want = want_all[k * n_line:k * n_line + len(lines)]
    r = subprocess.run([out, "0.9"], input=lines, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    assert r.stdout == want
