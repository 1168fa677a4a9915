"""Sharded verify_batch on ONE GPU: the ranks of a multi-GPU batch are played one after the other on the same device
(`verifier.verify_batch_shard` is what a rank runs; no NCCL), and the joint decision (`distributed.decide`) and the merged
state must equal the oracle's single-process result -- in particular when transactions of different shards touch the same
(account, asset): the reference threads `state` through the batch in order (src/tx/verify.rs:301-336,354-374; the
realistic_test, src/lib.rs:831-949, is this shape).  SURVEY.md 8e."""
import pytest

import oracle
import scenarios
from oracle import NATIVE

pytestmark = pytest.mark.gpu
OK, SIG, DECOMP, EQ, VAL, GENERIC, RANGE, TRANSCRIPT, FORMAT, NONCE, STATE, PARSE = range(12)
SEED = b"shard-test"


@pytest.fixture(scope="module")
def ctx():
    import xelis_he_b200 as xhe
    c = xhe.Ctx(0, party_capacity=16)
    yield c
    c.close()


def host_ledger(records, multisig=()):
    from xelis_he_b200.verifier import Ledger
    led = Ledger(); led.import_records(records)
    for pk, signers, th in multisig:
        led.set_multisig(pk, signers, th)
    return led


def run_sharded(ctx, blobs, records, world, mode, multisig=(), cuts=None):
    """every simulated rank verifies its shard against its own replica of the initial state; returns the joint verdict and
    the ledger obtained by applying the ranks' held-back updates in shard order (= what sync_and_commit builds on every rank)"""
    from xelis_he_b200 import distributed as xd, verifier
    n = len(blobs)
    bounds = cuts or [xd.shard_bounds(n, r, world) for r in range(world)]
    recs, handles = [], []
    for lo, hi in bounds:
        led = host_ledger(records, multisig)
        code, idx, s_enc, r_enc, tm = verifier.verify_batch_shard(ctx, blobs, led, lo, hi, seed=SEED, threads=2, fiat_shamir=mode)
        recs.append(xd.pack_local(code, idx, 0, s_enc, r_enc))
        handles.append(verifier.take_pending(ctx))
    verdict = xd.decide(recs, lambda encs: xd.sum_is_identity(ctx, encs))
    merged = host_ledger(records, multisig)
    for h in handles:
        if verdict[0] == OK and h:
            assert verifier.commit_taken(h, merged) == 0
        elif h:
            verifier.drop_taken(h)
    return verdict, merged


def check(ctx, blobs, oracle_ledger, records, multisig=(), worlds=(2, 3), modes=("host", "device", "fast"), cuts=None):
    ol = oracle_ledger.clone()
    want = oracle.verify_batch(blobs, ol)
    for world in worlds:
        for mode in modes:
            verdict, merged = run_sharded(ctx, blobs, records, world, mode, multisig, cuts)
            assert verdict == want, (world, mode, verdict, want)
            if want[0] == OK:
                assert merged.dump() == sorted(ol.dump()), (world, mode)
    return want


def _mut(blob, off, bit=1):
    b = bytearray(blob); b[off] ^= bit; return bytes(b)


def test_realistic_world_split_across_ranks(ctx):
    """src/lib.rs:831-949: tx2 spends what tx1 delivered.  With two ranks tx2 lands on rank 1."""
    w, txs, _ = scenarios.realistic_world()
    assert check(ctx, txs, w.ledger, w.records, worlds=(2,)) == (OK, -1)
    # the dependent transaction verified against the stale balance is exactly what must NOT happen: alone it is rejected
    from xelis_he_b200 import verifier
    assert verifier.verify_batch(ctx, [txs[1]], w.host_ledger(), seed=SEED)[0] != OK


def test_both_shards_credit_one_receiver_and_it_spends(ctx):
    w, txs = scenarios.shared_receiver_world(6)
    assert check(ctx, txs, w.ledger, w.records) == (OK, -1)
    # every cut of the batch gives the same result
    n = len(txs)
    for cut in range(1, n):
        assert check(ctx, txs, w.ledger, w.records, worlds=(2,), modes=("fast",), cuts=[(0, cut), (cut, n)]) == (OK, -1)


def test_one_sender_chain_across_ranks(ctx):
    """benches/tx.rs:153-186: every transaction of the batch comes from one sender, a length-T balance chain"""
    b = oracle.mint_chain(9, 12, 1)
    assert check(ctx, b.blobs, b.ledger(), b.ledger().dump(), worlds=(2, 3, 4)) == (OK, -1)


def test_independent_batch_and_bad_proofs_in_each_shard(ctx):
    b = oracle.mint_transfers(61, 12, 1, 1, threads=4)
    records = b.ledger().dump()
    assert check(ctx, b.blobs, b.ledger(), records) == (OK, -1)
    # a broken signature in the second shard: global index
    bad = list(b.blobs); bad[9] = _mut(bad[9], -1)
    assert check(ctx, bad, b.ledger(), records) == (SIG, 9)
    # one in each shard: the earliest transaction of the batch wins
    bad[2] = _mut(bad[2], 56)                                  # nonce
    assert check(ctx, bad, b.ledger(), records) == (NONCE, 2)


def test_resigned_bad_proofs_in_different_shards(ctx):
    """proofs that pass the signature and die in the MSM checks: the partial sums of the ranks must add up to a non-identity"""
    w = scenarios.World(b"shard-bad")
    accts = [w.account(b"s%d" % i, [(NATIVE, 1000)]) for i in range(6)]
    rcv = w.account(b"rcv", [(NATIVE, 0)])
    txs = [oracle.build_tx(kp, w.ledger, w.rng, fee=1, transfers=[(NATIVE, rcv.pk, 5 + i)], balances=[(NATIVE, 1000)]) for i, kp in enumerate(accts)]
    assert check(ctx, txs, w.ledger, w.records, worlds=(2,)) == (OK, -1)
    # validity proof z_x of tx 4 (second shard), re-signed: GenericProof, no transaction index
    bad = list(txs); bad[4] = oracle.resign(_mut(txs[4], 64 + 160 + 128), accts[4], w.rng)
    assert check(ctx, bad, w.ledger, w.records, worlds=(2, 3)) == (GENERIC, -1)
    # range proof t_x of tx 1 (first shard), re-signed: RangeProof
    rp0 = 64 + 324
    bad2 = list(txs); bad2[1] = oracle.resign(_mut(txs[1], rp0 + 128), accts[1], w.rng)
    assert check(ctx, bad2, w.ledger, w.records, worlds=(2, 3)) == (RANGE, -1)
    # both: the sigma check comes first
    bad3 = list(bad); bad3[1] = bad2[1]
    assert check(ctx, bad3, w.ledger, w.records, worlds=(2,)) == (GENERIC, -1)


def test_multisig_setup_in_an_earlier_shard(ctx):
    """src/lib.rs:254-612: the MultiSig transaction of shard 0 governs the co-signed spend in shard 1"""
    w, t, _ = scenarios.multisig_world()
    for spend, want in ((t["spend"], (OK, -1)), (t["spend_one"], None), (t["spend_none"], None)):
        txs = [t["setup"], spend]
        got = check(ctx, txs, w.ledger, w.records, worlds=(2,), modes=("host", "device", "fast"))
        if want:
            assert got == want
        else:
            assert got[0] != OK and got[1] == 1


def test_factors_differ_between_shards(ctx):
    """the batch factors absorb the index of the transaction in the WHOLE batch, so equal positions of two shards
    do not share a factor: forged proofs whose error terms cancel under equal factors are still rejected."""
    from xelis_he_b200 import distributed as xd, verifier
    b = oracle.mint_transfers(62, 8, 1, 1, threads=4)
    records = b.ledger().dump()
    recs = []
    for lo, hi in ((0, 4), (4, 8)):
        code, idx, s_enc, r_enc, _ = verifier.verify_batch_shard(ctx, b.blobs, host_ledger(records), lo, hi, seed=SEED, threads=2, fiat_shamir="device", deterministic=True)
        verifier.drop_taken(verifier.take_pending(ctx))
        recs.append((code, idx, s_enc, r_enc))
    assert all(r[0] == OK and r[2] == bytes(32) and r[3] == bytes(32) for r in recs)


def test_device_side_record_and_joint_decision(ctx):
    """xhe_batch_record_dev + xhe_shard_decide_dev (what bench.py times at N > 1): the records of two simulated ranks gathered on
    the device give the same decision as distributed.decide on the host-side records"""
    import torch
    from xelis_he_b200 import distributed as xd, verifier
    b = oracle.mint_transfers(63, 10, 1, 1, threads=4)
    records = b.ledger().dump()
    lib = ctx.lib
    for blobs, want in ((list(b.blobs), (OK, -1)), (b.blobs[:7] + [_mut(b.blobs[7], 64 + 160 + 128)] + b.blobs[8:], None)):
        gathered = torch.zeros(160, dtype=torch.uint8, device="cuda")
        host_recs = []
        for r, (lo, hi) in enumerate(((0, 5), (5, 10))):
            code, idx, s_enc, r_enc, _ = verifier.verify_batch_shard(ctx, blobs, host_ledger(records), lo, hi, seed=SEED, threads=2, fiat_shamir="fast")
            verifier.drop_taken(verifier.take_pending(ctx))
            host_recs.append(xd.pack_local(code, idx, 0, s_enc, r_enc))
            if want:      # honest shard: the device can build the record itself from the resident batch
                assert lib.xhe_batch_record_dev(ctx.p, gathered.data_ptr() + 80 * r) == 0
                ctx.sync()
                assert bytes(gathered[80 * r:80 * r + 80].cpu().numpy())[:76] == host_recs[-1][:76]
            else:
                gathered[80 * r:80 * r + 80] = torch.frombuffer(bytearray(host_recs[-1]), dtype=torch.uint8).cuda()
        out = torch.zeros(16, dtype=torch.uint8, device="cuda")
        assert lib.xhe_shard_decide_dev(ctx.p, gathered.data_ptr(), 2, out.data_ptr()) == 0
        ctx.sync()
        raw = bytes(out.cpu().numpy())
        got = (int.from_bytes(raw[:4], "little", signed=True), int.from_bytes(raw[8:16], "little", signed=True))
        assert got == xd.decide(host_recs, lambda encs: xd.sum_is_identity(ctx, encs))
        if want:
            assert got == want
