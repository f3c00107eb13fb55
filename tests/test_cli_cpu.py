"""The C++ host (`metabuli-b200 classify`) end to end WITHOUT a GPU: the real CLI binary runs against a test double of the
C-ABI library (tests/host/stub_backend.cpp: the same entry points, answered by the oracle), so the host's own code — streaming
FASTA/FASTQ(.gz) reader with read-ahead, batch pipeline, paired files read side by side, --mask / --mask-host, row formatter,
report writer, flag handling — is compared byte for byte with the files the reference binary wrote.  The GPU suite runs the same
command lines against the real library (test_gpu_fixtures.py, test_gpu_synth.py)."""
import gzip
import os
import shutil
import subprocess

import pytest

import synth_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "metabuli_b200", "_lib", "metabuli-b200")
BUILD = os.path.join(ROOT, "tests", "host", "_build", "cli_stub")


@pytest.fixture(scope="module")
def cli():
    """A directory with a copy of the CLI and the test double under the real library's name (the CLI's rpath is $ORIGIN)."""
    if not os.path.exists(EXE):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "libmetabuli_b200.so")
    srcs = [os.path.join(ROOT, "tests", "host", "stub_backend.cpp"), os.path.join(ROOT, "oracle", "mbl_oracle.cpp")]
    deps = srcs + [os.path.join(ROOT, "oracle", "mbl_oracle.hpp"), os.path.join(ROOT, "include", "metabuli_b200.h"),
                   os.path.join(ROOT, "metabuli_b200", "csrc", "host", "tantan_mask.hpp")]
    if not os.path.exists(lib) or any(os.path.getmtime(lib) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-o", lib] + srcs +
                              ["-lz", "-lpthread"])
    exe = os.path.join(BUILD, "metabuli-b200")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(EXE):
        shutil.copy2(EXE, exe)
    return exe


def _run(exe, args, db_dir, out_dir, timeout=600):
    env = dict(os.environ, MBL_STUB_DB_DIR=db_dir)
    r = subprocess.run([exe, "classify"] + args + [db_dir, out_dir, "job"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-3000:]
    return (open(os.path.join(out_dir, "job_classifications.tsv"), "rb").read(), open(os.path.join(out_dir, "job_report.tsv"), "rb").read(), r.stdout)


@pytest.mark.parametrize("db,mode,batch,ext", [("in", "pe", 0, "fna.gz"), ("in", "se", 1500, "fna.gz"), ("ex", "pe", 777, "fq.gz"), ("ex", "se", 0, "fq.gz")])
def test_reference_fixture(cli, db, mode, batch, ext, fixtures_dir, golden_dir, tmp_path):
    reads = [os.path.join(fixtures_dir, "reads", f"ERR9594652_5000_{k}.{ext}") for k in ((1, 2) if mode == "pe" else (1,))]
    args = ["--seq-mode", "2" if mode == "pe" else "1", "--threads", "4"] + (["--batch-reads", str(batch)] if batch else []) + reads
    tsv, report, log = _run(cli, args, os.path.join(fixtures_dir, f"db_{db}"), str(tmp_path))
    assert tsv == gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_{mode}_classifications.tsv.gz"), "rb").read()
    assert report == gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_{mode}_report.tsv.gz"), "rb").read()
    assert "Total read count : 5000" in log


def _write_case(name, tmp_path, plain_fastq=False):
    sdb, reads, seq_mode = synth_cases.build(name)
    db_dir = str(tmp_path / "db")
    sdb.write(db_dir)
    files = []
    for k in range(len(reads) // 2):
        p = str(tmp_path / f"r{k + 1}.{'fq' if plain_fastq else 'fna'}")
        if plain_fastq:
            b, o = reads[2 * k], reads[2 * k + 1]
            with open(p, "wb") as f:
                for i in range(o.size - 1):
                    s = bytes(b[int(o[i]):int(o[i + 1])])
                    f.write(b"@r%d extra words\n%s\n+\n%s\n" % (i, s, b"@" * len(s)))        # qualities that look like headers
        else:
            synth_cases.write_fasta(p, reads[2 * k], reads[2 * k + 1])
        files.append(p)
    return db_dir, files, seq_mode


@pytest.mark.parametrize("name,extra", [("mask_se", []), ("mask_se", ["--mask-host", "1"]), ("mask_pe", []), ("mask_pe", ["--mask-host", "1"])])
def test_masked_queries(cli, name, extra, golden_dir, tmp_path):
    """--mask 1 through the host: with --mask-host 1 the reader thread masks (mbl_mask_reads), otherwise mbl_config.mask_mode
    asks the library (here: the double, which masks on upload like K0 does)."""
    db_dir, files, seq_mode = _write_case(name, tmp_path)
    mask, prob = synth_cases.mask_flags(name)
    args = ["--seq-mode", str(seq_mode), "--threads", "3", "--batch-reads", "700", "--mask", "1", "--mask-prob", str(prob)] + extra + files
    tsv, report, _ = _run(cli, args, db_dir, str(tmp_path))
    assert tsv == gzip.open(os.path.join(golden_dir, "synth", name + ".tsv.gz"), "rb").read()
    assert report == gzip.open(os.path.join(golden_dir, "synth", name + ".report.gz"), "rb").read()


@pytest.mark.parametrize("name", ["ragged_pe", "flags_se", "lineage_se", "sync_pe", "long", "format1_pe", "acc_lvl1_se", "redund_se"])
def test_synthetic_cases_through_the_host(cli, name, golden_dir, tmp_path):
    """Ragged paired mates from plain FASTQ (qualities starting with '@'), non-default thresholds, --lineage 1, a syncmer database."""
    db_dir, files, seq_mode = _write_case(name, tmp_path, plain_fastq=(name == "ragged_pe"))
    args = ["--seq-mode", str(seq_mode), "--threads", "4", "--batch-reads", "997"]
    for k, v in synth_cases.FLAGS.get(name, {}).items():
        args += [k, str(v)]
    tsv, report, _ = _run(cli, args + files, db_dir, str(tmp_path))
    assert tsv == gzip.open(os.path.join(golden_dir, "synth", name + ".tsv.gz"), "rb").read()
    assert report == gzip.open(os.path.join(golden_dir, "synth", name + ".report.gz"), "rb").read()


def test_unequal_mate_files_are_an_error(cli, tmp_path):
    db_dir, files, _ = _write_case("multi_pe", tmp_path)
    txt = open(files[1]).read().split(">")
    open(files[1], "w").write(">".join(txt[:-1]))                      # mate 2 lacks the last record
    env = dict(os.environ, MBL_STUB_DB_DIR=db_dir)
    r = subprocess.run([cli, "classify", "--seq-mode", "2", "--threads", "2"] + files + [db_dir, str(tmp_path), "job"], capture_output=True, text=True,
                       timeout=300, env=env)
    assert r.returncode != 0 and "The number of reads in the two files are not equal." in (r.stdout + r.stderr)


@pytest.mark.parametrize("devices,batch", [("0,1,2", 333), ("0,0", 1250)])
def test_replica_workers_keep_the_read_order(cli, devices, batch, fixtures_dir, golden_dir, tmp_path):
    """--devices a,b,c: one worker thread per context, batches handed to whichever is free, rows written in input order (the
    double ignores the device ordinal, so this is the host's scheduling and ordering logic alone)."""
    reads = [os.path.join(fixtures_dir, "reads", f"ERR9594652_5000_{k}.fna.gz") for k in (1, 2)]
    args = ["--seq-mode", "2", "--threads", "4", "--devices", devices, "--batch-reads", str(batch)] + reads
    tsv, report, log = _run(cli, args, os.path.join(fixtures_dir, "db_in"), str(tmp_path))
    assert tsv == gzip.open(os.path.join(golden_dir, "ref_tsv", "in_pe_classifications.tsv.gz"), "rb").read()
    assert report == gzip.open(os.path.join(golden_dir, "ref_tsv", "in_pe_report.tsv.gz"), "rb").read()
    assert "completed on %d GPU(s)" % len(devices.split(",")) in log


def test_damaged_gzip_input_is_an_error(cli, fixtures_dir, tmp_path):
    """A truncated or corrupted .gz stops the run with the decoder's message (CRC / length of every member are checked); no partial
    TSV is reported as success."""
    src = os.path.join(fixtures_dir, "reads", "ERR9594652_5000_1.fna.gz")
    data = open(src, "rb").read()
    db_dir = os.path.join(fixtures_dir, "db_in")
    env = dict(os.environ, MBL_STUB_DB_DIR=db_dir)
    for label, blob in (("cut", data[: len(data) // 2]), ("flip", data[:5000] + bytes([data[5000] ^ 0x20]) + data[5001:])):
        p = tmp_path / (label + ".fna.gz")
        p.write_bytes(blob)
        r = subprocess.run([cli, "classify", "--seq-mode", "1", "--threads", "2", str(p), db_dir, str(tmp_path), label], capture_output=True, text=True,
                           timeout=300, env=env)
        assert r.returncode != 0 and "gzip:" in (r.stdout + r.stderr), (label, r.stdout[-500:], r.stderr[-500:])


def test_large_gzip_input_goes_through_the_parallel_decoder(cli, fixtures_dir, golden_dir, tmp_path):
    """Two .gz mate files big enough (> 2 x 2 MB compressed) for the reader to decode each of them with several threads: the
    fixture's pairs x 24 -> the reference's rows x 24."""
    files = []
    for k in (1, 2):
        raw = gzip.open(os.path.join(fixtures_dir, "reads", f"ERR9594652_5000_{k}.fna.gz"), "rb").read()
        p = tmp_path / f"big_{k}.fna.gz"
        with gzip.open(p, "wb", compresslevel=1) as f:
            for _ in range(24):
                f.write(raw)
        assert os.path.getsize(p) > 5 << 20
        files.append(str(p))
    tsv, _, log = _run(cli, ["--seq-mode", "2", "--threads", "6", "--batch-reads", "50000"] + files, os.path.join(fixtures_dir, "db_in"), str(tmp_path), timeout=1200)
    gold = gzip.open(os.path.join(golden_dir, "ref_tsv", "in_pe_classifications.tsv.gz"), "rb").read()
    nl = gold.index(b"\n") + 1
    assert tsv == gold[:nl] + gold[nl:] * 24
    assert "Total read count : 120000" in log


def test_taxonomy_path_is_ignored_like_in_the_reference(cli, fixtures_dir, golden_dir, tmp_path):
    """loadTaxonomy (common.cpp:50-86) reads --taxonomy-path only when <dbdir>/taxonomyDB is missing; with the file present the
    reference binary writes the same TSV even for a path that does not exist (checked here in round 2).  Same behaviour: accepted,
    no effect; a database without taxonomyDB is an error that says why."""
    reads = [os.path.join(fixtures_dir, "reads", "ERR9594652_5000_1.fna.gz")]
    tsv, _, _ = _run(cli, ["--seq-mode", "1", "--threads", "3", "--taxonomy-path", "/no/such/dir"] + reads, os.path.join(fixtures_dir, "db_in"), str(tmp_path))
    assert tsv == gzip.open(os.path.join(golden_dir, "ref_tsv", "in_se_classifications.tsv.gz"), "rb").read()
    bare = tmp_path / "db_without_taxonomy"
    bare.mkdir()
    for f in ("diffIdx", "info", "split", "taxID_list", "db.parameters"):
        src = os.path.join(fixtures_dir, "db_in", f)
        if os.path.exists(src):
            shutil.copy(src, bare / f)
    r = subprocess.run([cli, "classify", "--seq-mode", "1"] + reads + [str(bare), str(tmp_path), "x"], capture_output=True, text=True,
                       env=dict(os.environ, MBL_STUB_DB_DIR=str(bare)))
    assert r.returncode != 0 and "taxonomyDB is NOT found" in (r.stdout + r.stderr)


@pytest.mark.parametrize("devices,batch,db,mode", [("0,1", 1300, "in", "pe"), ("0,1,2", 999, "ex", "se"), ("0,1,2,3", 0, "in", "se")])
def test_index_sharded_orchestration(cli, devices, batch, db, mode, fixtures_dir, golden_dir, tmp_path):
    """--index-sharded 1: the host cuts every batch over the ranks and runs extract -> exchange -> match -> exchange -> score in
    lock step (barriers, count tables through host memory, receive buffers grown on demand, results stitched back in read order).
    The double answers the phases on host memory; what is checked is the host's orchestration: TSV and report of the reference."""
    reads = [os.path.join(fixtures_dir, "reads", f"ERR9594652_5000_{k}.fna.gz") for k in ((1, 2) if mode == "pe" else (1,))]
    args = ["--seq-mode", "2" if mode == "pe" else "1", "--threads", "4", "--devices", devices, "--index-sharded", "1"]
    args += (["--batch-reads", str(batch)] if batch else []) + reads
    tsv, report, log = _run(cli, args, os.path.join(fixtures_dir, f"db_{db}"), str(tmp_path))
    assert tsv == gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_{mode}_classifications.tsv.gz"), "rb").read()
    assert report == gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_{mode}_report.tsv.gz"), "rb").read()
    assert "index sharded" in log


def test_more_ranks_than_reads_in_a_batch(cli, fixtures_dir, golden_dir, tmp_path):
    """Batches of three reads over four ranks (a rank without reads in every round) and of one read over three replicas."""
    raw = gzip.open(os.path.join(fixtures_dir, "reads", "ERR9594652_5000_1.fna.gz"), "rb").read().split(b"\n")
    small = tmp_path / "small.fna"
    small.write_bytes(b"\n".join(raw[:46]) + b"\n")                     # 23 records
    want = b"\n".join(gzip.open(os.path.join(golden_dir, "ref_tsv", "in_se_classifications.tsv.gz"), "rb").read().split(b"\n")[:24]) + b"\n"
    db = os.path.join(fixtures_dir, "db_in")
    for extra in (["--devices", "0,1,2,3", "--index-sharded", "1", "--batch-reads", "3"], ["--devices", "0,1,2", "--batch-reads", "1"]):
        tsv, _, _ = _run(cli, ["--seq-mode", "1", "--threads", "2"] + extra + [str(small)], db, str(tmp_path))
        assert tsv == want, extra
