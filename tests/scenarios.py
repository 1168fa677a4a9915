"""Shared scenario builders for the verify_batch parity tests: the reference's own test cases (src/lib.rs:254-1093,
src/proofs.rs:378-447) replayed through the oracle's prover, plus re-signed bad-proof cases the reference never tests
(SURVEY.md 4).  Each scenario = (name, [tx blobs], [(pk, asset, ct, nonce)], multisig [(pk, [signers], threshold)])."""
import oracle
from oracle import NATIVE, Keypair, Rng

ASSET55 = bytes([55]) * 32


class World:
    def __init__(self, tag):
        self.rng = Rng(b"scenario-" + tag)
        self.ledger = oracle.Ledger()
        self.records = []
        self.multisig = []

    def account(self, name, balances):
        kp = Keypair.derive(name)
        for asset, amount in balances:
            ct = kp.encrypt(amount, self.rng)
            self.ledger.set_balance(kp.pk, asset, ct)
            self.records.append((kp.pk, asset, ct))
        self.ledger.set_nonce(kp.pk, 0)
        return kp

    def set_multisig(self, pk, signers, threshold):
        self.ledger.set_multisig(pk, signers, threshold)
        self.multisig.append((pk, signers, threshold))

    def host_ledger(self):
        from xelis_he_b200.verifier import Ledger
        led = Ledger()
        led.import_records(self.records)
        for pk, signers, threshold in self.multisig:
            led.set_multisig(pk, signers, threshold)
        return led


def burn_world():   # src/lib.rs:614-657 test_burn
    w = World(b"burn")
    alice = w.account(b"alice", [(NATIVE, 100)])
    tx = oracle.build_tx(alice, w.ledger, w.rng, fee=1, burn=(NATIVE, 10), balances=[(NATIVE, 100)])
    return w, [tx], alice


def burn_non_native_world():   # src/lib.rs:659-703
    w = World(b"burn55")
    alice = w.account(b"alice", [(NATIVE, 1), (ASSET55, 50)])
    tx = oracle.build_tx(alice, w.ledger, w.rng, fee=1, burn=(ASSET55, 50), balances=[(NATIVE, 1), (ASSET55, 50)])
    return w, [tx], alice


def realistic_world():   # src/lib.rs:831-949 realistic_test: two dependent multi-destination, multi-asset txs
    w = World(b"realistic")
    bob = w.account(b"bob", [(NATIVE, 100), (ASSET55, 2)])
    alice = w.account(b"alice", [(NATIVE, 0), (ASSET55, 0)])
    eve = w.account(b"eve", [(NATIVE, 52), (ASSET55, 0)])
    tx1 = oracle.build_tx(bob, w.ledger, w.rng, fee=1, transfers=[(NATIVE, alice.pk, 52), (NATIVE, eve.pk, 4), (ASSET55, eve.pk, 2)],
                          balances=[(NATIVE, 100), (ASSET55, 2)])
    after = w.ledger.clone()
    assert oracle.apply_without_verify(tx1, after) == 0
    tx2 = oracle.build_tx(alice, after, w.rng, fee=1, transfers=[(NATIVE, eve.pk, 30)], balances=[(NATIVE, 52)])
    return w, [tx1, tx2], (bob, alice, eve)


def multisig_world():   # src/lib.rs:254-612: setup tx, then a tx from the multisig account co-signed by threshold signers
    w = World(b"multisig")
    alice = w.account(b"alice", [(NATIVE, 100)])
    bob = w.account(b"bob", [(NATIVE, 0)])
    charlie = w.account(b"charlie", [(NATIVE, 0)])
    dave = w.account(b"dave", [(NATIVE, 0)])
    setup = oracle.build_tx(alice, w.ledger, w.rng, fee=1, multisig_setup=([charlie.pk, dave.pk], 2), balances=[(NATIVE, 100)])
    after = w.ledger.clone()
    assert oracle.apply_without_verify(setup, after) == 0
    spend = oracle.build_tx(alice, after, w.rng, fee=1, transfers=[(NATIVE, bob.pk, 10)], balances=[(NATIVE, 99)],
                            cosigners=[(0, charlie), (1, dave)])
    spend_one = oracle.build_tx(alice, after, w.rng, fee=1, transfers=[(NATIVE, bob.pk, 10)], balances=[(NATIVE, 99)], cosigners=[(0, charlie)])
    spend_dup = oracle.build_tx(alice, after, w.rng, fee=1, transfers=[(NATIVE, bob.pk, 10)], balances=[(NATIVE, 99)], cosigners=[(0, charlie), (0, charlie)])
    spend_wrong = oracle.build_tx(alice, after, w.rng, fee=1, transfers=[(NATIVE, bob.pk, 10)], balances=[(NATIVE, 99)], cosigners=[(0, dave), (1, charlie)])
    spend_none = oracle.build_tx(alice, after, w.rng, fee=1, transfers=[(NATIVE, bob.pk, 10)], balances=[(NATIVE, 99)])
    return w, dict(setup=setup, spend=spend, spend_one=spend_one, spend_dup=spend_dup, spend_wrong=spend_wrong, spend_none=spend_none), (alice, bob, charlie, dave)


def mixed_types_world(n=12):   # config 5 flavour: transfers (k 1..4, a 1..2), burn, call-contract, deploy, multisig setup
    w = World(b"mixed")
    accts = [w.account(b"acct%d" % i, [(NATIVE, 10**6), (ASSET55, 10**6)]) for i in range(n)]
    # receivers never send in this batch: a sender's proof is bound to its balance at build time (cf. realistic_test)
    rcv = [w.account(b"rcv%d" % i, [(NATIVE, 5), (ASSET55, 0)]) for i in range(3)]
    txs = []
    for i, kp in enumerate(accts):
        dest = rcv[i % 3].pk
        kind = i % 6
        if kind == 0:
            txs.append(oracle.build_tx(kp, w.ledger, w.rng, fee=3, transfers=[(NATIVE, dest, 7)], balances=[(NATIVE, 10**6)]))
        elif kind == 1:
            txs.append(oracle.build_tx(kp, w.ledger, w.rng, fee=3, transfers=[(NATIVE, dest, 1), (ASSET55, dest, 2), (NATIVE, rcv[(i + 1) % 3].pk, 3)],
                                       balances=[(NATIVE, 10**6), (ASSET55, 10**6)]))
        elif kind == 2:
            txs.append(oracle.build_tx(kp, w.ledger, w.rng, fee=2, burn=(ASSET55, 500), balances=[(NATIVE, 10**6), (ASSET55, 10**6)]))
        elif kind == 3:
            txs.append(oracle.build_tx(kp, w.ledger, w.rng, fee=2, call=(bytes([9]) * 32, [(NATIVE, 40)], [(b"method", b"swap")]), balances=[(NATIVE, 10**6)]))
        elif kind == 4:
            txs.append(oracle.build_tx(kp, w.ledger, w.rng, fee=1, deploy=b"contract code bytes", balances=[(NATIVE, 10**6)]))
        else:
            txs.append(oracle.build_tx(kp, w.ledger, w.rng, fee=1, multisig_setup=([accts[(i + 1) % n].pk, accts[(i + 2) % n].pk], 1), balances=[(NATIVE, 10**6)]))
    return w, txs, accts


def transfer_with_extra_data_world():   # src/lib.rs:951-1029: memo bytes are covered by the signature only
    w = World(b"extra")
    bob = w.account(b"bob", [(NATIVE, 100)])
    alice = w.account(b"alice", [(NATIVE, 0)])
    memo = b"\x01\x02\x03 encrypted memo bytes" + bytes(64)   # cipher || sender_handle || receiver_handle (opaque to the verifier)
    tx = oracle.build_tx(bob, w.ledger, w.rng, fee=1, transfers=[(NATIVE, alice.pk, 5, memo)], balances=[(NATIVE, 100)])
    return w, [tx], (bob, alice)


def shared_receiver_world(n=6):   # several senders (in different shards) credit ONE receiver, which then spends the lot
    w = World(b"shared-rcv")
    senders = [w.account(b"snd%d" % i, [(NATIVE, 1000)]) for i in range(n - 1)]
    rcv = w.account(b"rcv", [(NATIVE, 10)])
    sink = w.account(b"sink", [(NATIVE, 0)])
    txs, state, total = [], w.ledger.clone(), 10
    for i, kp in enumerate(senders):
        tx = oracle.build_tx(kp, w.ledger, w.rng, fee=1, transfers=[(NATIVE, rcv.pk, 20 + i)], balances=[(NATIVE, 1000)])
        assert oracle.apply_without_verify(tx, state) == 0
        txs.append(tx); total += 20 + i
    # the receiver spends everything it has by now: valid only against the balance AFTER all the credits
    txs.append(oracle.build_tx(rcv, state, w.rng, fee=2, transfers=[(NATIVE, sink.pk, total - 2)], balances=[(NATIVE, total)]))
    return w, txs
