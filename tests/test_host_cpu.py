"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/xhe.h declares, fails loudly without
a device, and its host-side building blocks (Merlin, SHA3, BLAKE3, wide reduction, to_bytes) agree with public known
answers and with the oracle."""
import ctypes as C
import hashlib
import os
import re

import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import xelis_he_b200 as xhe
    return xhe.load_library()


def test_library_exports_every_declared_symbol(lib):
    for header, pattern, least in (("xhe.h", r"\b(xhe_[a-z0-9_]+)\s*\(", 20), ("xhe_host.h", r"\b(xheh_[a-z0-9_]+)\s*\(", 20)):
        hdr = open(os.path.join(ROOT, "include", header)).read()
        names = sorted(set(re.findall(pattern, hdr)))
        assert len(names) >= least, header
        missing = [n for n in names if not hasattr(lib, n)]
        assert not missing, (header, missing)


def test_rust_sys_crate_matches_the_header(tmp_path):
    """xhe-sys/ is source-only here (no Rust toolchain): pin what can be pinned without one -- every extern it declares is
    in include/xhe.h, and the struct sizes its layout test asserts are the C compiler's sizes for the header's structs."""
    import subprocess
    rs = open(os.path.join(ROOT, "xhe-sys", "src", "lib.rs")).read()
    hdr = open(os.path.join(ROOT, "include", "xhe.h")).read()
    declared = set(re.findall(r"\b(xhe_[a-z0-9_]+)\s*\(", hdr))
    fns = set(re.findall(r"pub fn (xhe_[a-z0-9_]+)\(", rs))
    assert len(fns) >= 20 and not (fns - declared), fns - declared
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "xhe.h"\nint main(void){printf("%zu %zu\\n", sizeof(xhe_batch), sizeof(xhe_verdict));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    c_batch, c_verdict = subprocess.check_output([str(exe)]).split()
    assert re.search(r"size_of::<xhe_batch>\(\), (\d+)\)", rs).group(1) == c_batch.decode()
    assert re.search(r"size_of::<xhe_verdict>\(\), (\d+)\)", rs).group(1) == c_verdict.decode()
    # field order of the Rust structs = field order of the C structs
    def c_fields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            for part in decl.split(","):
                m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*(\[\d+\])?\s*$", part.strip())
                if m and part.strip():
                    out.append(m.group(1))
        return out
    def rs_fields(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, rs, re.S).group(1)
        return re.findall(r"pub ([a-z0-9_]+):", body)
    for name in ("xhe_batch", "xhe_verdict"):
        assert c_fields(name) == rs_fields(name), name


def test_build_is_keyed_on_a_source_hash():
    from xelis_he_b200 import build
    assert not build.needs_build() and open(build.HASH_FILE).read().strip() == build.source_hash()


def test_no_cpu_fallback_without_a_device():
    import torch
    import xelis_he_b200 as xhe
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(xhe.XheError):
        xhe.Ctx(0, party_capacity=2)


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under xelis_he_b200/ may import, include, link or dlopen it"""
    bad = re.compile(r"^\s*(import|from)\s+oracle\b|#include\s+\"[^\"]*oracle|liboracle|dlopen[^\n]*oracle", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "xelis_he_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(src), os.path.join(dirpath, f)


def test_host_merlin_and_hashes(lib):
    out = C.create_string_buffer(32)
    lib.xheh_merlin_test(b"test protocol", b"some label", b"some data", C.c_size_t(9), b"challenge", out, C.c_size_t(32))
    assert out.raw.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    blake3 = pytest.importorskip("blake3")
    for n in (0, 1, 71, 72, 73, 135, 136, 137, 1023, 1024, 1025, 1500, 3073, 5000):
        msg = bytes((i * 7 + 3) & 0xFF for i in range(n))
        o64 = C.create_string_buffer(64); lib.xheh_sha3_512(msg, C.c_size_t(n), o64)
        assert o64.raw == hashlib.sha3_512(msg).digest()
        o200 = C.create_string_buffer(200); lib.xheh_shake256(msg, C.c_size_t(n), o200, C.c_size_t(200))
        assert o200.raw == hashlib.shake_256(msg).digest(200)
        o32 = C.create_string_buffer(32); lib.xheh_blake3(msg, C.c_size_t(n), o32)
        assert o32.raw == blake3.blake3(msg).digest()


def test_host_wide_reduction(lib):
    L = 2**252 + 27742317777372353535851937790883648493
    stream = hashlib.shake_256(b"wide").digest(64 * 500)
    cases = [stream[64 * i:64 * i + 64] for i in range(500)] + [bytes(64), b"\xff" * 64, (L - 1).to_bytes(64, "little"), L.to_bytes(64, "little"), ((L << 256) + L - 1).to_bytes(64, "little")]
    for x in cases:
        o = C.create_string_buffer(32); lib.xheh_reduce_wide(x, o)
        assert int.from_bytes(o.raw, "little") == int.from_bytes(x, "little") % L
        assert o.raw == oracle.sc_reduce_wide(x)


def test_constant_2_64_G(lib):
    """the encoding the host adds as a term when fee + amount reaches 2^64 (get_sender_output_ct adds Scalars)"""
    import ctypes as C
    import oracle
    out = C.create_string_buffer(32)
    lib.xheh_const_g_2_64(out)
    G = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")      # RFC 9496 generator
    assert out.raw == oracle.msm((2 ** 64).to_bytes(32, "little"), G)


def test_host_to_bytes_matches_oracle(lib):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import scenarios
    blobs = []
    for f in (scenarios.burn_world, scenarios.realistic_world, scenarios.transfer_with_extra_data_world):
        blobs += f()[1]
    blobs += list(scenarios.multisig_world()[1].values()) + scenarios.mixed_types_world(6)[1]
    lib.xheh_tx_to_bytes.restype = C.c_int32
    for b in blobs:
        want, msi = oracle.tx_to_bytes(b)
        out = C.create_string_buffer(len(b) + 64); n = C.c_size_t(0); m = C.c_size_t(0)
        assert lib.xheh_tx_to_bytes(b, C.c_size_t(len(b)), out, C.c_size_t(len(out)), C.byref(n), C.byref(m)) == 0
        assert out.raw[:n.value] == want and m.value == msi
    # framing errors are rejected by both parsers
    for bad in (blobs[0][:-1], blobs[0] + b"\x00", b"\x00" * 100):
        out = C.create_string_buffer(4096); n = C.c_size_t(0); m = C.c_size_t(0)
        assert lib.xheh_tx_to_bytes(bad, C.c_size_t(len(bad)), out, C.c_size_t(4096), C.byref(n), C.byref(m)) == 11
        assert oracle.tx_to_bytes(bad) is None


def test_key_digest_index_finds_the_same_dependencies_as_the_blob_scan(lib):
    """Sharded batches: the earlier transactions a shard depends on (src/tx/verify.rs:301-374) found through the
    batch's key-digest index must be exactly those the scan of the blobs finds -- dependent worlds and an independent batch."""
    import scenarios
    from xelis_he_b200 import verifier
    ms = scenarios.multisig_world()[1]
    cases = [list(scenarios.realistic_world()[1]), list(scenarios.shared_receiver_world(6)[1]), [ms["setup"], ms["spend"]], list(scenarios.mixed_types_world(12)[1])] + [list(oracle.mint_chain(5, 24, 1).blobs), list(oracle.mint_transfers(9, 40, 1, 2, threads=4).blobs)]
    found_dependency = False
    for blobs in cases:
        n = len(blobs)
        plain, indexed = verifier.prepare_blobs(blobs), verifier.prepare_blobs(blobs, index=True)
        assert indexed.index and indexed.index_bytes >= 4 * (n + 1)
        for lo in range(0, n + 1):
            for hi in sorted({lo, min(n, lo + 1), min(n, lo + 3), n}):
                want = verifier.shard_dependencies(plain, lo, hi)
                assert all(i < lo for i in want) and want == sorted(want)
                assert verifier.shard_dependencies(indexed, lo, hi, threads=3) == want, (n, lo, hi)
                found_dependency |= bool(want)
    assert found_dependency
    indep = list(oracle.mint_transfers(9, 40, 1, 2, threads=4).blobs)
    assert verifier.shard_dependencies(verifier.prepare_blobs(indep, index=True), 20, 40) == []


def test_fast_path_host_phases_run_without_a_device(lib):
    """The host half of the fast path (header pass, state walk with the cache-resident chain table, cross-shard dependency
    lookup through the key-digest index, staging) is plain C++: the diagnostics flag of xheh_verify_batch_ex / _shard_ix runs
    it without a device.  Nothing is verified in that mode, so the call reports an error code, never XHE_OK."""
    import ctypes as C
    import scenarios
    from xelis_he_b200 import verifier
    lib.xheh_verify_batch_shard_ix.restype = C.c_int32
    lib.xheh_verify_batch_shard_ix.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int, C.c_uint32,
                                               C.POINTER(C.c_long), C.c_void_p, C.c_void_p, C.c_void_p]
    b = oracle.mint_transfers(21, 64, 1, 2, threads=4)
    chain = oracle.mint_chain(22, 16, 1)
    w, txs = scenarios.shared_receiver_world(6)
    for blobs, records in ((list(b.blobs), b.ledger().dump()), (list(chain.blobs), chain.ledger().dump()), (list(txs), w.ledger.dump())):
        led = verifier.Ledger(); led.import_records(records)
        bl = verifier.prepare_blobs(blobs, index=True)
        n = len(blobs)
        for lo, hi in ((0, n), (n // 2, n), (n - 1, n)):
            fi = C.c_long(-1); tm = (C.c_double * 7)(); part = C.create_string_buffer(64)
            rc = lib.xheh_verify_batch_shard_ix(C.c_void_p(4096), led.ptr, bl.ptrs, bl.lens, n, lo, hi, b"dry", 3, 2, 4 | 16, C.byref(fi), tm, part, bl.index)
            assert rc == -1, rc                      # XHE_E_ARG: a dry run never claims a verdict
            assert tm[5] >= tm[0] >= 0 and tm[1] >= 0
