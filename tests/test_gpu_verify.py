"""GPU parity for Transaction::verify_batch / verify / apply_without_verify through the host layer + C ABI, against the
CPU oracle: identical accept/reject verdicts (code and first failing tx) on honest and tampered batches, and byte-identical
updated balances.  Scenarios replay the reference's tests (src/lib.rs:254-1093) and add re-signed bad-proof cases."""
import pytest

import oracle
import scenarios
from oracle import NATIVE

pytestmark = pytest.mark.gpu
OK, SIG, DECOMP, EQ, VAL, GENERIC, RANGE, TRANSCRIPT, FORMAT, NONCE, STATE, PARSE = range(12)
SEED = b"test-batch-factors"


@pytest.fixture(scope="module")
def ctx():
    import xelis_he_b200 as xhe
    c = xhe.Ctx(0, party_capacity=16)
    yield c
    c.close()


def _fresh(batch):
    from xelis_he_b200 import verifier
    hl = verifier.Ledger(); hl.import_records(batch.ledger().dump())
    return hl


def both(ctx, world, txs, expect=None):
    """run oracle and device on clones of the same ledger; assert identical verdicts (+ states on accept)"""
    from xelis_he_b200 import verifier
    ol = world.ledger.clone()
    want = oracle.verify_batch(txs, ol)
    hl = world.host_ledger()
    code, idx, tm = verifier.verify_batch(ctx, txs, hl, seed=SEED, threads=4)
    assert (code, idx) == want, (code, idx, want)
    hl_dev = world.host_ledger()                      # same batch with device-side Fiat-Shamir: same verdict, same state
    code_d, idx_d, _ = verifier.verify_batch(ctx, txs, hl_dev, seed=SEED, threads=4, fiat_shamir="device")
    assert (code_d, idx_d) == want, ("device fiat-shamir", code_d, idx_d, want)
    assert hl_dev.dump() == hl.dump()
    hl_fast = world.host_ledger()                     # optimistic device-layout path (falls back to the exact path on any failure)
    code_f, idx_f, tm_f = verifier.verify_batch(ctx, txs, hl_fast, seed=SEED, threads=4, fiat_shamir="fast")
    assert (code_f, idx_f) == want, ("fast path", code_f, idx_f, want)
    assert hl_fast.dump() == hl.dump()
    long_transcripts = any(t[1] == 0 and int.from_bytes(t[4:8], "little") >= 32 for t in txs)      # few transactions with hundreds of transfers: Merlin runs on the host
    if code == OK and txs and not any(t[1] == 4 or t[3] != 0xFF for t in txs) and not world.multisig and not long_transcripts:
        assert tm_f["fast_path"], "an honest non-multisig batch must be decided by the fast path"
    if expect is not None:
        assert code == expect
    if code == OK:
        assert hl.dump() == sorted(ol.dump())
    return code, idx


def test_minted_batches_accept(ctx):
    for a, k, T in ((1, 1, 40), (1, 3, 12), (2, 6, 6)):
        b = oracle.mint_transfers(7 + a + k, T, a, k, threads=8)
        from xelis_he_b200 import verifier
        hl = verifier.Ledger(); hl.import_records(b.ledger().dump())
        ol = b.ledger()
        assert oracle.verify_batch(b.blobs, ol) == (OK, -1)
        code, idx, tm = verifier.verify_batch(ctx, b.blobs, hl, seed=SEED)
        assert (code, idx) == (OK, -1)
        assert hl.dump() == sorted(ol.dump())
        hf = verifier.Ledger(); hf.import_records(b.ledger().dump())
        code, idx, tm = verifier.verify_batch(ctx, b.blobs, hf, seed=SEED, fiat_shamir="fast")
        assert (code, idx) == (OK, -1) and tm["fast_path"] and hf.dump() == sorted(ol.dump())


def test_mixed_party_sizes_in_one_batch(ctx):
    from xelis_he_b200 import verifier
    bs = [oracle.mint_transfers(31, 5, 1, 1), oracle.mint_transfers(32, 4, 2, 6), oracle.mint_transfers(33, 3, 1, 3)]
    blobs = [x for b in bs for x in b.blobs]
    recs = [r for b in bs for r in b.ledger().dump()]
    hl = verifier.Ledger(); hl.import_records(recs)
    ol = oracle.Ledger()
    for pk, asset, ct in recs:
        ol.set_balance(pk, asset, ct); ol.set_nonce(pk, 0)
    assert oracle.verify_batch(blobs, ol) == (OK, -1)
    assert verifier.verify_batch(ctx, blobs, hl, seed=SEED)[:2] == (OK, -1)
    assert hl.dump() == sorted(ol.dump())


def test_255_transfer_transaction_m256():
    """benches/tx.rs:109 `n_tx_bench(c, 255)`: one sender, 255 transfers, a 256-party aggregated range proof (lg = 14),
    alone and inside a batch with small transactions (mixed m in one range MSM); a re-signed bad range proof is rejected
    with RangeProof on both sides; a context created for fewer parties refuses loudly instead of mis-verifying."""
    import xelis_he_b200 as xhe
    from xelis_he_b200 import verifier
    w = scenarios.World(b"m256")
    bob = w.account(b"bob", [(NATIVE, 10000000)]); alice = w.account(b"alice", [(NATIVE, 0)]); carol = w.account(b"carol", [(NATIVE, 500)])
    big = oracle.build_tx(bob, w.ledger, w.rng, fee=3, transfers=[(NATIVE, alice.pk, 1)] * 255, balances=[(NATIVE, 10000000)])
    small = oracle.build_tx(carol, w.ledger, w.rng, fee=1, transfers=[(NATIVE, alice.pk, 7)], balances=[(NATIVE, 500)])
    c = xhe.Ctx(0, party_capacity=256)
    try:
        assert both(c, w, [big]) == (OK, -1)
        assert both(c, w, [small, big]) == (OK, -1)
        rp0 = 64 + 324 * 255
        bad = oracle.resign(_mut(big, rp0 + 128), bob, w.rng)              # range proof t_x
        assert both(c, w, [bad]) == (RANGE, -1)
        assert both(c, w, [small, bad]) == (RANGE, -1)
    finally:
        c.close()
    c8 = xhe.Ctx(0, party_capacity=8)
    try:
        with pytest.raises(Exception):
            verifier.verify_batch(c8, [big], w.host_ledger(), seed=SEED)
    finally:
        c8.close()


def test_single_sender_chain(ctx):   # benches/tx.rs:129-186 shape: balances chained through one account
    from xelis_he_b200 import verifier
    b = oracle.mint_chain(5, 24, 1)
    hl = verifier.Ledger(); hl.import_records(b.ledger().dump())
    ol = b.ledger()
    assert oracle.verify_batch(b.blobs, ol) == (OK, -1)
    assert verifier.verify_batch(ctx, b.blobs, hl, seed=SEED)[:2] == (OK, -1)
    assert hl.dump() == sorted(ol.dump())
    # a chain verified out of order must fail exactly where the oracle says (eq proof bound to the running balance)
    swapped = [b.blobs[1], b.blobs[0]] + b.blobs[2:]
    hl2 = verifier.Ledger(); hl2.import_records(b.ledger().dump())
    assert verifier.verify_batch(ctx, swapped, hl2, seed=SEED)[:2] == oracle.verify_batch(swapped, b.ledger()) == (GENERIC, -1)


def test_reference_scenarios_accept(ctx):
    for f in (scenarios.burn_world, scenarios.burn_non_native_world, scenarios.realistic_world, scenarios.transfer_with_extra_data_world):
        w, txs, _ = f()
        both(ctx, w, txs, OK)
    w, txs, _ = scenarios.mixed_types_world()
    both(ctx, w, txs, OK)
    both(ctx, w, [], OK)            # empty batch


def test_multisig_scenarios(ctx):
    w, d, (alice, bob, charlie, dave) = scenarios.multisig_world()
    both(ctx, w, [d["setup"]], OK)
    both(ctx, w, [d["setup"], d["spend"]], OK)
    assert both(ctx, w, [d["setup"], d["spend_one"]]) == (FORMAT, 1)     # signature count != threshold
    assert both(ctx, w, [d["setup"], d["spend_dup"]]) == (FORMAT, 1)     # duplicate signer index
    assert both(ctx, w, [d["setup"], d["spend_wrong"]]) == (SIG, 1)      # signatures by the wrong keys
    assert both(ctx, w, [d["setup"], d["spend_none"]]) == (FORMAT, 1)    # multisig account, tx without multisig
    assert both(ctx, w, [d["spend"]]) == (FORMAT, 0)                     # multisig in tx, none in state


def _mut(blob, off, xor=1):
    b = bytearray(blob); b[off] ^= xor; return bytes(b)


def test_tampered_transactions_reject_like_reference(ctx):
    # src/lib.rs:705-829: every mutation changes to_bytes() and dies at the signature (or nonce / format) check
    w, txs, alice = scenarios.burn_world()
    tx = txs[0]
    assert both(ctx, w, [_mut(tx, len(tx) - 40)])[0] == SIG              # signature scalar e
    assert both(ctx, w, [_mut(tx, 64 + 32)])[0] == SIG                   # burn amount
    assert both(ctx, w, [_mut(tx, 64)])[0] == FORMAT                     # burn asset no longer has a commitment
    assert both(ctx, w, [_mut(tx, 48)])[0] == SIG                        # fee
    assert both(ctx, w, [_mut(tx, 56)])[0] == NONCE                      # nonce
    cleared = bytearray(tx); cleared[2] = 0
    assert both(ctx, w, [bytes(cleared)])[0] in (PARSE, FORMAT)          # commitments cleared: framing breaks first in our wire format
    w2, txs2, _ = scenarios.realistic_world()
    assert both(ctx, w2, [txs2[0], _mut(txs2[1], 60)]) == (NONCE, 1)     # first failing tx index is reported
    assert both(ctx, w2, [txs2[1], txs2[0]])[0] == GENERIC               # tx2 before tx1: stale balance, sigma MSM fails


def test_resigned_bad_proofs_hit_the_msm_checks(ctx):
    """The reference's own tests never reach GenericProof / RangeProof: re-sign after mutating the proofs."""
    w = scenarios.World(b"resign")
    bob = w.account(b"bob", [(NATIVE, 1000)]); alice = w.account(b"alice", [(NATIVE, 0)])
    tx = oracle.build_tx(bob, w.ledger, w.rng, fee=1, transfers=[(NATIVE, alice.pk, 5)], balances=[(NATIVE, 1000)])
    # layout: hdr 64 | transfer: asset32 dest32 C32 Ds32 Dr32 proof160 (Y0 Y1 Y2 z_r z_x) extra4 | range proof | commitments
    t0 = 64
    bad_zx = oracle.resign(_mut(tx, t0 + 160 + 128), bob, w.rng)          # validity proof z_x
    assert both(ctx, w, [bad_zx]) == (GENERIC, -1)
    rp0 = t0 + 324
    bad_tx = oracle.resign(_mut(tx, rp0 + 128), bob, w.rng)               # range proof t_x
    assert both(ctx, w, [bad_tx]) == (RANGE, -1)
    bad_L = oracle.resign(_mut(tx, rp0 + 224 + 3), bob, w.rng)            # range proof L_0: most likely not a valid point
    assert both(ctx, w, [bad_L])[0] == RANGE
    y0_zero = bytearray(tx); y0_zero[t0 + 160:t0 + 192] = bytes(32)
    assert both(ctx, w, [oracle.resign(bytes(y0_zero), bob, w.rng)]) == (TRANSCRIPT, 0)   # identity Y_0
    y1_bad = bytearray(tx); y1_bad[t0 + 192:t0 + 224] = (1).to_bytes(32, "little")       # negative s: not a point
    assert both(ctx, w, [oracle.resign(bytes(y1_bad), bob, w.rng)]) == (VAL, 0)
    c_bad = bytearray(tx); c_bad[t0 + 64:t0 + 96] = (1).to_bytes(32, "little")           # amount commitment not a point
    assert both(ctx, w, [oracle.resign(bytes(c_bad), bob, w.rng)]) == (DECOMP, 0)
    sc0 = rp0 + 32 * (9 + 2 * 7)
    eq_bad = oracle.resign(_mut(tx, sc0 + 64 + 96 + 5), bob, w.rng)       # eq proof z_s
    assert both(ctx, w, [eq_bad]) == (GENERIC, -1)
    a_zero = bytearray(tx); a_zero[rp0:rp0 + 32] = bytes(32)              # range proof A = identity encoding
    assert both(ctx, w, [oracle.resign(bytes(a_zero), bob, w.rng)])[0] == RANGE
    # a good tx followed by a bad one: still rejected; good alone accepted
    both(ctx, w, [tx], OK)


def test_apply_without_verify_matches_oracle(ctx):
    from xelis_he_b200 import verifier
    w, txs, _ = scenarios.realistic_world()
    ol = w.ledger.clone()
    for t in txs:
        assert oracle.apply_without_verify(t, ol) == 0
    hl = w.host_ledger()
    assert verifier.apply_without_verify(ctx, txs, hl) == 0
    assert hl.dump() == sorted(ol.dump())


def test_output_ciphertexts_match_oracle(ctx):
    """BlockchainVerificationState::set_output_ciphertext (src/tx/verify.rs:60-66; called at 339-340 and 582): a state that
    asks for them receives get_sender_output_ct(source, asset) per transaction, byte-identical (compressed) to the oracle's,
    through every placement of the host/device split and through apply_without_verify; balances are unaffected."""
    from xelis_he_b200 import verifier
    worlds = [scenarios.realistic_world()[:2], scenarios.burn_world()[:2], (lambda r: (r[0], r[1]))(scenarios.mixed_types_world(12))]
    for w, txs in worlds:
        ol = w.ledger.clone().record_outputs()
        assert oracle.verify_batch(txs, ol) == (OK, -1)
        want_out, want_bal = sorted(ol.dump_outputs()), sorted(ol.dump())
        assert want_out, "scenario produced no output ciphertexts"
        for mode in ("host", "device", "fast"):
            hl = w.host_ledger(); hl.record_outputs()
            assert verifier.verify_batch(ctx, txs, hl, seed=SEED, threads=4, fiat_shamir=mode)[:2] == (OK, -1)
            assert hl.dump_outputs() == want_out, mode
            assert hl.dump() == want_bal, mode
        quiet = w.host_ledger()                        # a state that does not ask gets none (and the same balances)
        assert verifier.verify_batch(ctx, txs, quiet, seed=SEED, threads=4, fiat_shamir="fast")[:2] == (OK, -1)
        assert quiet.dump_outputs() == [] and quiet.dump() == want_bal
    # minted a2k6 batch through the fast path, and apply_without_verify
    b = oracle.mint_transfers(77, 6, 2, 6, threads=8)
    ol = b.ledger().record_outputs(); assert oracle.verify_batch(b.blobs, ol) == (OK, -1)
    hl = verifier.Ledger(); hl.import_records(b.ledger().dump()); hl.record_outputs()
    code, idx, tm = verifier.verify_batch(ctx, b.blobs, hl, seed=SEED, fiat_shamir="fast")
    assert (code, idx) == (OK, -1) and tm["fast_path"]
    assert hl.dump_outputs() == sorted(ol.dump_outputs()) and hl.dump() == sorted(ol.dump())
    w, txs, _ = scenarios.realistic_world()
    ol = w.ledger.clone().record_outputs()
    for t in txs:
        assert oracle.apply_without_verify(t, ol) == 0
    hl = w.host_ledger(); hl.record_outputs()
    assert verifier.apply_without_verify(ctx, txs, hl) == 0
    assert hl.dump_outputs() == sorted(ol.dump_outputs()) and hl.dump() == sorted(ol.dump())


def test_plain_amounts_add_as_scalars_not_u64(ctx):
    """get_sender_output_ct adds fee and burn / contract amounts as Scalars (src/tx/verify.rs:107-144, src/elgamal.rs:353-377):
    fee = 2^64 - 1 plus a burn of 5 is 2^64 + 4, not 4.  A prover who wraps the sum gets a transaction the reference rejects;
    verdict, balances (apply_without_verify) and output ciphertext must follow the reference, not the wrapped value."""
    from xelis_he_b200 import verifier
    w = scenarios.World(b"amount-overflow")
    alice = w.account(b"alice", [(NATIVE, 100)])
    tx = oracle.build_tx(alice, w.ledger, w.rng, fee=2**64 - 1, burn=(NATIVE, 5), balances=[(NATIVE, 100)])
    code, _ = both(ctx, w, [tx])
    assert code != OK
    ol = w.ledger.clone().record_outputs()
    assert oracle.apply_without_verify(tx, ol) == 0
    hl = w.host_ledger(); hl.record_outputs()
    assert verifier.apply_without_verify(ctx, [tx], hl) == 0
    assert hl.dump() == sorted(ol.dump())
    assert hl.dump_outputs() == sorted(ol.dump_outputs())


def test_device_fiat_shamir_matches_host(ctx):
    """both modes derive the same challenges and batch factors: the partial MSM encodings of a (deliberately invalid) shard
    are byte-identical, not just the verdicts"""
    from xelis_he_b200 import verifier
    b = oracle.mint_transfers(91, 9, 2, 3, threads=8)
    blobs = list(b.blobs)
    blobs[4] = _mut(blobs[4], 64 + 160 + 128)      # corrupt a validity-proof response: the sigma partial is no longer the identity
    outs = []
    for mode in ("host", "device"):
        hl = verifier.Ledger(); hl.import_records(b.ledger().dump())
        outs.append(verifier.verify_batch_partial(ctx, blobs, hl, seed=SEED, threads=3, fiat_shamir=mode, deterministic=True)[:4])
    assert outs[0] == outs[1] and outs[0][0] == SIG and outs[0][1] == 4 and outs[0][2] != bytes(32)
    # outside the replay flag the factors are unpredictable -- the same call twice gives different partial sums
    a1 = verifier.verify_batch_partial(ctx, blobs, _fresh(b), seed=SEED, fiat_shamir="device")[2]
    a2 = verifier.verify_batch_partial(ctx, blobs, _fresh(b), seed=SEED, fiat_shamir="device")[2]
    assert a1 != a2 and a1 != outs[0][2]
    swapped = [b.blobs[1], b.blobs[0]] + b.blobs[2:]          # valid txs: partials are the identity in both modes
    for mode in ("host", "device", "fast"):
        hl = verifier.Ledger(); hl.import_records(b.ledger().dump())
        r = verifier.verify_batch_partial(ctx, swapped, hl, seed=SEED, fiat_shamir=mode)
        assert r[:4] == (0, -1, bytes(32), bytes(32))
        assert r[4]["fast_path"] == (mode == "fast")


def test_verdict_partials_combine(ctx):
    """multi-GPU algebra on one device: the sigma / range partial sums of two halves add up to the identity iff each does."""
    import ctypes as C
    # exercised through xhe_combine_partials with the identity and P + (-P)
    ident = (C.c_uint8 * 128)(); ident[32] = 1; ident[64] = 1      # X=0, Y=1, Z=1, T=0 packed
    lib = ctx.lib
    lib.xhe_combine_partials.restype = C.c_int32
    lib.xhe_combine_partials.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int32)]
    out = (C.c_uint8 * 32)(); flag = C.c_int32(0)
    assert lib.xhe_combine_partials(ctx.p, bytes(ident) * 3, 3, out, C.byref(flag)) == 0
    assert flag.value == 1 and bytes(out) == bytes(32)


def test_fixed_base_and_generic_static_msm_agree(ctx):
    """The MSM over the static range-proof generators runs from a fixed-base table on contexts of up to 64 parties and
    through the generic Pippenger pipeline above: for the same shard (same seed, so the same batch factors) the partial
    range / sigma encodings must be byte-identical -- including a shard whose range sum is NOT the identity."""
    import xelis_he_b200 as xhe
    from xelis_he_b200 import verifier
    b = oracle.mint_transfers(91, 12, 2, 6, threads=8)               # m = 8
    big = xhe.Ctx(0, party_capacity=128)                              # > 64: generic pipeline
    try:
        for blobs in (list(b.blobs), list(b.blobs[:5])):
            outs = []
            for c in (ctx, big):
                for mode in ("host", "fast"):
                    hl = verifier.Ledger(); hl.import_records(b.ledger().dump())
                    code, idx, sig, rng, _ = verifier.verify_batch_partial(c, blobs, hl, seed=SEED, fiat_shamir=mode)
                    assert (code, idx) == (OK, -1)
                    outs.append((sig, rng))
            assert outs[0] == outs[2] and outs[1] == outs[3], "fixed-base and generic static MSM disagree"
        # a deliberately broken range proof (re-signing is not needed for the partial sums: signatures are checked apart)
        bad = list(b.blobs[:3]); raw = bytearray(bad[1])
        k, rp0 = 6, 64 + 324 * 6
        raw[rp0 + 128] ^= 1; bad[1] = bytes(raw)
        res = []
        for c in (ctx, big):
            hl = verifier.Ledger(); hl.import_records(b.ledger().dump())
            res.append(verifier.verify_batch_partial(c, bad, hl, seed=SEED, fiat_shamir="host", deterministic=True)[:4])
        assert res[0] == res[1] and res[0][3] != bytes(32)
    finally:
        big.close()


def test_large_batch_size_independent_properties(ctx):
    """2,000-transfer batch (the bench workload's shape at a size the oracle still checks in seconds): accept through the
    fast path, every updated ciphertext equal to the oracle's apply_without_verify, and a tampered transaction deep inside
    the batch rejected at the index and with the code the oracle reports."""
    from xelis_he_b200 import verifier
    T = 2000
    b = oracle.mint_transfers(91, T, 1, 1, threads=16)
    hl = verifier.Ledger(); hl.import_records(b.ledger().dump())
    code, idx, tm = verifier.verify_batch(ctx, b.blobs, hl, seed=SEED, fiat_shamir="fast")
    assert (code, idx) == (OK, -1) and tm["fast_path"]
    ol = b.ledger()
    for blob in b.blobs:
        assert oracle.apply_without_verify(blob, ol) == 0
    assert hl.dump() == sorted(ol.dump())
    # idempotence of the verdict under a different batch-factor seed, and the host-transcript split agrees
    hl2 = verifier.Ledger(); hl2.import_records(b.ledger().dump())
    assert verifier.verify_batch(ctx, b.blobs, hl2, seed=b"another seed", fiat_shamir="host")[:2] == (OK, -1)
    assert hl2.dump() == hl.dump()
    bad = list(b.blobs); victim = 1617
    t = bytearray(bad[victim]); t[-1] ^= 0x01; bad[victim] = bytes(t)           # last byte of the signature
    hl3 = verifier.Ledger(); hl3.import_records(b.ledger().dump())
    got = verifier.verify_batch(ctx, bad, hl3, seed=SEED, fiat_shamir="fast")
    assert got[:2] == oracle.verify_batch(bad[:victim + 1], b.slice(victim + 1).ledger()) == (1, victim)
    assert got[2]["fast_path"]       # decided by the fast path + the exact verdict of that ONE transaction, not by a re-run of the batch


def test_long_single_sender_chain(ctx):
    """C3's one-sender variant: 256 transactions spending from one account, so every source balance is the output of the
    previous transaction (pointer-jumping prefix over a chain of length T)."""
    from xelis_he_b200 import verifier
    b = oracle.mint_chain(77, 256, 1)
    ol = b.ledger()
    assert oracle.verify_batch(b.blobs, ol) == (OK, -1)
    for mode in ("fast", "host"):
        hl = verifier.Ledger(); hl.import_records(b.ledger().dump())
        assert verifier.verify_batch(ctx, b.blobs, hl, seed=SEED, fiat_shamir=mode)[:2] == (OK, -1)
        assert hl.dump() == sorted(ol.dump())


def test_mixed_batch_config5_flavour(ctx):
    """Config 5 at a size the oracle's (sequential) builder mints in seconds: 96 transactions mixing multi-destination and
    multi-asset transfers, burns, contract calls, deploys and multisig set-ups; accept run, then one tampered transaction per
    class in the middle of the batch (the first failing index and code must match the oracle)."""
    w, txs, _ = scenarios.mixed_types_world(96)
    assert both(ctx, w, txs, OK) == (OK, -1)
    for victim, off in ((37, -1), (50, 48), (63, 56)):                # signature byte, fee, nonce
        bad = list(txs); t = bytearray(bad[victim]); t[off if off >= 0 else len(t) - 1] ^= 0x01; bad[victim] = bytes(t)
        code, idx = both(ctx, w, bad)
        assert idx == victim and code in (SIG, NONCE)


def test_joint_msm_failures_keep_the_reference_precedence(ctx):
    """The accepting path sums the sigma and the range check in ONE MSM; a sum that is not the identity is re-run as two MSMs.
    A batch in which BOTH checks fail (in different transactions, in either order) must report GenericProof like the reference
    (sigma check first, src/tx/verify.rs:500-514); a failing range proof alone RangeProof; and a batch in which a sigma error
    and a range error are made to look alike must still not be accepted."""
    w = scenarios.World(b"joint-msm")
    accts = [w.account(b"acct%d" % i, [(NATIVE, 1000)]) for i in range(6)]
    sink = w.account(b"sink", [(NATIVE, 0)])
    txs = [oracle.build_tx(a, w.ledger, w.rng, fee=1, transfers=[(NATIVE, sink.pk, 5 + i)], balances=[(NATIVE, 1000)]) for i, a in enumerate(accts)]
    t0, rp0 = 64, 64 + 324
    bad_sigma = oracle.resign(_mut(txs[1], t0 + 160 + 128), accts[1], w.rng)       # validity proof z_x of tx 1
    bad_range = oracle.resign(_mut(txs[4], rp0 + 128), accts[4], w.rng)            # range proof t_x of tx 4
    bad_range_early = oracle.resign(_mut(txs[0], rp0 + 128), accts[0], w.rng)
    assert both(ctx, w, txs) == (OK, -1)
    assert both(ctx, w, txs[:1] + [bad_sigma] + txs[2:4] + [bad_range] + txs[5:]) == (GENERIC, -1)
    assert both(ctx, w, [bad_range_early] + [bad_sigma] + txs[2:]) == (GENERIC, -1)            # the range error comes first in the batch: still sigma first
    assert both(ctx, w, txs[:4] + [bad_range] + txs[5:]) == (RANGE, -1)
    assert both(ctx, w, txs[:1] + [bad_sigma] + txs[2:]) == (GENERIC, -1)


def test_full_size_batch_with_a_tampered_transaction(ctx):
    """BASELINE's batch size (10,000 transfers): a signature flipped deep inside the batch is rejected with the code and index the
    oracle reports for the prefix that ends there, and the honest batch is accepted with every balance equal to the oracle's."""
    from xelis_he_b200 import verifier
    T = 10000
    b = oracle.mint_transfers(93, T, 1, 1, threads=16)
    records = b.ledger().dump()
    hl = verifier.Ledger(); hl.import_records(records)
    code, idx, tm = verifier.verify_batch(ctx, b.blobs, hl, seed=SEED, fiat_shamir="fast")
    assert (code, idx) == (OK, -1) and tm["fast_path"]
    ol = b.ledger()
    for blob in b.blobs:
        assert oracle.apply_without_verify(blob, ol) == 0
    assert hl.dump() == sorted(ol.dump())
    victim = 8765
    bad = list(b.blobs); t = bytearray(bad[victim]); t[-1] ^= 0x01; bad[victim] = bytes(t)
    hl2 = verifier.Ledger(); hl2.import_records(records)
    got = verifier.verify_batch(ctx, bad, hl2, seed=SEED, fiat_shamir="fast")
    assert got[:2] == oracle.verify_batch(bad[:victim + 1], b.slice(victim + 1).ledger()) == (1, victim)
    assert hl2.dump() == sorted(records)                      # a rejected batch leaves the state untouched
