"""GPU parity tests (through the C ABI) for the arithmetic layer and the per-point kernels, against the CPU oracle and
Python big integers.  Bit-exact: every comparison is on bytes."""
import hashlib
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493


@pytest.fixture(scope="module")
def ctx():
    import xelis_he_b200 as xhe
    c = xhe.Ctx(0, party_capacity=8)
    yield c
    c.close()


def _words(vals):
    return np.array([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for v in vals], dtype=np.uint32)


def _ints(arr):
    return [sum(int(arr[r, i]) << (32 * i) for i in range(8)) for r in range(arr.shape[0])]


def test_field_ops_vs_bigint(ctx):
    rng = random.Random(7)
    edge = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2 * P, 2 * P + 37, 2**256 - 1, 2**256 - 38, 2**255, 2**255 - 1, 2**32 - 1, (2**256 - 1) ^ (2**128)]
    a = edge * len(edge) + [rng.getrandbits(256) for _ in range(4096)]
    b = [y for y in edge for _ in edge] + [rng.getrandbits(256) for _ in range(4096)]
    A, B = _words(a), _words(b)
    for op, f in ((0, lambda x, y: x + y), (1, lambda x, y: x - y), (2, lambda x, y: x * y), (3, lambda x, y: x * x), (7, lambda x, y: -x)):
        got = _ints(ctx.selftest_fe(op, A, B))
        for x, y, g in zip(a, b, got):
            assert g < 2**256 and g % P == f(x, y) % P, (op, hex(x), hex(y))
    got = _ints(ctx.selftest_fe(4, A, B))
    assert got == [x % P for x in a]
    sub = [x for x in a[:512] if x % P]
    got = _ints(ctx.selftest_fe(5, _words(sub), _words(sub)))
    assert all(g * x % P == 1 for g, x in zip(got, sub))
    got = _ints(ctx.selftest_fe(6, _words(sub), _words(sub)))
    assert all(g % P == pow(x, (P - 5) // 8, P) for g, x in zip(got, sub))


def test_scalar_ops_vs_bigint(ctx):
    rng = random.Random(11)
    edge = [0, 1, L - 1, L, L + 1, 2**256 - 1, 2**252]
    a = edge * len(edge) + [rng.getrandbits(256) for _ in range(2048)]
    b = [y for y in edge for _ in edge] + [rng.getrandbits(256) for _ in range(2048)]
    A, B = _words(a), _words(b)
    assert _ints(ctx.selftest_fe(8, A, B)) == [x * y % L for x, y in zip(a, b)]
    assert _ints(ctx.selftest_fe(9, A, B)) == [(x + (y << 256)) % L for x, y in zip(a, b)]
    nz = [x for x in a if x % L]
    got = _ints(ctx.selftest_fe(10, _words(nz), _words(nz)))
    assert all(g * x % L == 1 for g, x in zip(got, nz))


def _sample_encodings(n_valid=600, n_random=600):
    import oracle
    rnd = hashlib.shake_256(b"gpu-basic").digest(64 * n_valid)
    valid = [oracle.from_uniform(rnd[64 * i:64 * i + 64]) for i in range(n_valid)]
    random_enc = hashlib.shake_256(b"random-encodings-gpu").digest(32 * n_random)
    special = [bytes(32), (P).to_bytes(32, "little"), (1).to_bytes(32, "little"), b"\xff" * 32, (2).to_bytes(32, "little"),
               (P - 1).to_bytes(32, "little"), bytes(31) + b"\x80"]
    return b"".join(valid) + random_enc + b"".join(special)


def test_decompress_compress_parity(ctx):
    import oracle
    enc = _sample_encodings()
    n = len(enc) // 32
    ok_ref, xy_ref = oracle.decode_batch(enc, want_xy=True)
    ok, xy = ctx.decompress(enc, want_xy=True)
    assert ok == ok_ref
    for i in range(n):
        if ok[i]:
            assert xy[64 * i:64 * i + 64] == xy_ref[64 * i:64 * i + 64], i
    good = [i for i in range(n) if ok[i]]
    assert 600 < len(good) < n
    xy_good = b"".join(xy[64 * i:64 * i + 64] for i in good)
    assert ctx.compress(xy_good) == b"".join(enc[32 * i:32 * i + 32] for i in good)
    assert ctx.decompress(b"") == b""


def test_from_uniform_parity(ctx):
    import oracle
    u = hashlib.shake_256(b"uniform-gpu").digest(64 * 500) + bytes(64) + b"\xff" * 64
    got = ctx.from_uniform(u)
    for i in range(len(u) // 64):
        assert got[32 * i:32 * i + 32] == oracle.from_uniform(u[64 * i:64 * i + 64]), i


def test_ct_update_parity(ctx):
    import oracle
    rnd = hashlib.shake_256(b"ct-gpu").digest(64 * 1200)
    pts = [oracle.from_uniform(rnd[64 * i:64 * i + 64]) for i in range(1200)]
    n = 300
    bal = b"".join(pts[0:2 * n]); delta = b"".join(pts[2 * n:4 * n])
    # edge cases: identity balance, delta == balance (sub -> identity), invalid encodings
    bal = bytes(64) + pts[5] + pts[6] + bal[128:]
    delta = pts[1] + pts[2] + pts[5] + pts[6] + delta[128:]
    bal = bal[:64 * 10] + b"\x01" + bal[64 * 10 + 1:]        # negative s: invalid
    sub = bytes([i & 1 for i in range(n)])
    sub = bytes([0, 1]) + sub[2:]
    want, ok_ref = oracle.ct_update(bal, delta, sub)
    got, ok = ctx.ct_update(bal, delta, sub)
    assert ok == ok_ref and ok[10] == 0 and sum(ok) == n - 1
    assert got == want
    assert got[64:128] == bytes(64)   # P - P = identity encoding


def _dev_points(ctx, torch, n, seed):
    """n pseudo-random valid ristretto255 encodings on the device (one-way map of seeded bytes)"""
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    uni = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda", generator=g)
    enc = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    assert ctx.lib.xhe_from_uniform_dev(ctx.p, uni.data_ptr(), n, enc.data_ptr()) == 0
    return enc


def test_ct_update_resident_parity(ctx):
    """config 4, resident layout (the HBM-bound kernel): planar extended balances +/- planar affine-Niels deltas, updated
    in place; re-encoded results must equal the compressed-I/O kernel and the CPU oracle byte for byte."""
    import torch
    import oracle
    lib = ctx.lib
    n = 2048; npts = 2 * n
    bal_enc = _dev_points(ctx, torch, npts, 11); del_enc = _dev_points(ctx, torch, npts, 12)
    sub = (torch.arange(n, device="cuda") % 3 == 0).to(torch.uint8)
    aff = torch.empty((npts, 16), dtype=torch.int32, device="cuda"); ok = torch.empty((npts,), dtype=torch.uint8, device="cuda")
    niels = torch.empty((npts, 24), dtype=torch.int32, device="cuda"); aff_d = torch.empty((npts, 16), dtype=torch.int32, device="cuda")
    assert lib.xhe_decompress_dev(ctx.p, bal_enc.data_ptr(), npts, aff.data_ptr(), None, ok.data_ptr()) == 0
    assert lib.xhe_decompress_dev(ctx.p, del_enc.data_ptr(), npts, aff_d.data_ptr(), niels.data_ptr(), ok.data_ptr()) == 0
    torch.cuda.synchronize()
    a = aff.cpu().numpy().view(np.uint32)
    xs, ys = _ints(a[:, :8]), _ints(a[:, 8:])
    T = _words([x * y % P for x, y in zip(xs, ys)])
    one = np.zeros((npts, 8), dtype=np.uint32); one[:, 0] = 1
    planar = np.stack([a[:, :8], a[:, 8:], one, T])                                  # [X|Y|Z|T][2n][8]
    bal = torch.from_numpy(planar.view(np.int32).copy()).cuda()
    dn = niels.view(npts, 3, 8).permute(1, 0, 2).contiguous()                        # [ypx|ymx|t2d][2n][8]
    assert lib.xhe_ct_update_resident_dev(ctx.p, bal.data_ptr(), dn.data_ptr(), sub.data_ptr(), n) == 0
    ext = bal.permute(1, 0, 2).contiguous()                                           # back to X,Y,Z,T per point
    got = torch.empty((npts, 32), dtype=torch.uint8, device="cuda")
    assert lib.xhe_compress_dev(ctx.p, ext.data_ptr(), npts, got.data_ptr()) == 0
    out = torch.empty((n, 64), dtype=torch.uint8, device="cuda"); okb = torch.empty((n,), dtype=torch.uint8, device="cuda")
    assert lib.xhe_ct_update_dev(ctx.p, bal_enc.data_ptr(), del_enc.data_ptr(), sub.data_ptr(), n, out.data_ptr(), okb.data_ptr()) == 0
    torch.cuda.synchronize()
    got_b = bytes(got.cpu().numpy().tobytes()); out_b = bytes(out.cpu().numpy().tobytes())
    assert int(okb.sum()) == n and got_b == out_b
    want, ok_ref = oracle.ct_update(bytes(bal_enc.cpu().numpy().tobytes()), bytes(del_enc.cpu().numpy().tobytes()), bytes(sub.cpu().numpy().tobytes()))
    assert all(ok_ref) and got_b == want


def test_ct_update_full_size_round_trip(ctx):
    """config 4 at BASELINE's size (1,048,576 accounts, compressed I/O): (bal +/- delta) -/+ delta == bal byte for byte
    (encodings are canonical), every account valid, and a 1,024-account sample equals the oracle."""
    import torch
    import oracle
    lib = ctx.lib
    n = 1 << 20
    bal = _dev_points(ctx, torch, 2 * n, 21).view(n, 64); delta = _dev_points(ctx, torch, 2 * n, 22).view(n, 64)
    g = torch.Generator(device="cuda"); g.manual_seed(23)
    sub = torch.randint(0, 2, (n,), dtype=torch.uint8, device="cuda", generator=g)
    out = torch.empty((n, 64), dtype=torch.uint8, device="cuda"); back = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
    ok = torch.empty((n,), dtype=torch.uint8, device="cuda")
    assert lib.xhe_ct_update_dev(ctx.p, bal.data_ptr(), delta.data_ptr(), sub.data_ptr(), n, out.data_ptr(), ok.data_ptr()) == 0
    torch.cuda.synchronize(); assert int(ok.sum()) == n
    inv = 1 - sub
    assert lib.xhe_ct_update_dev(ctx.p, out.data_ptr(), delta.data_ptr(), inv.data_ptr(), n, back.data_ptr(), ok.data_ptr()) == 0
    torch.cuda.synchronize(); assert int(ok.sum()) == n
    assert torch.equal(back, bal) and not torch.equal(out, bal)
    idx = torch.arange(0, n, n // 1024, device="cuda")[:1024]
    want, ok_ref = oracle.ct_update(bytes(bal[idx].cpu().numpy().tobytes()), bytes(delta[idx].cpu().numpy().tobytes()), bytes(sub[idx].cpu().numpy().tobytes()))
    assert all(ok_ref) and bytes(out[idx].cpu().numpy().tobytes()) == want


def test_int_peak_reports(ctx):
    rates = [ctx.int_peak(w) for w in range(3)]
    print("int peak (inst/s): IMAD.lo %.3e IMAD.HI %.3e IMAD.WIDE %.3e" % tuple(rates))
    assert all(r > 1e11 for r in rates)


def test_sum_encodings_parity(ctx):
    """cross-rank combination of partial MSM results: sum of encodings == oracle point additions; identity and bad input"""
    import oracle
    rnd = hashlib.shake_256(b"sum-enc").digest(64 * 40)
    pts = [oracle.from_uniform(rnd[64 * i:64 * i + 64]) for i in range(40)]
    ident = bytes(32)
    for n in (0, 1, 2, 8, 33, 40):
        want = ident
        for p in pts[:n]:
            want = oracle.point_add(want, p)
        enc, is_id = ctx.sum_encodings(b"".join(pts[:n]))
        assert enc == want and is_id == (want == ident)
    a = pts[0]; neg_a = oracle.point_add(ident, a, sub=True)
    assert ctx.sum_encodings(a + pts[1] + neg_a + oracle.point_add(ident, pts[1], sub=True)) == (ident, True)
    bad = b"\xff" * 32                       # non-canonical field element: not a valid encoding
    assert ctx.sum_encodings(a + bad + neg_a)[1] is False


def test_warp_cooperative_arithmetic_vs_bigint(ctx):
    """csrc/oct.cuh (the Horner chain of the MSM): field elements spread over 8 lanes, points over a warp -- products, sums and
    differences against Python integers (edge values around p, 2p and 2^256 exercise the carry look-ahead and the 2^255 = 19
    fold), doubling and complete addition against the affine twisted-Edwards law (a = -1)."""
    rng = random.Random(17)
    edge = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2 * P, 2 * P + 37, 2**256 - 1, 2**256 - 38, 2**255, 2**255 - 1, 2**32 - 1, (2**256 - 1) ^ (2**128), 2**256 - 2**32, (2**256 - 1) ^ 0xFFFFFFFF]
    a = edge * len(edge) + [rng.getrandbits(256) for _ in range(4099)]
    b = [y for y in edge for _ in edge] + [rng.getrandbits(256) for _ in range(4099)]
    A, B = _words(a), _words(b)
    for op, f in ((0, lambda x, y: x * y), (1, lambda x, y: x + y), (2, lambda x, y: x - y)):
        got = _ints(ctx.selftest_oct(op, A, B))
        for x, y, g in zip(a, b, got):
            assert g < 2**256 and g % P == f(x, y) % P, (op, hex(x), hex(y))
    # points: random multiples of the base point in extended coordinates with random Z, plus the identity and a 2-torsion point
    D = (-121665 * pow(121666, P - 2, P)) % P

    def aff_add(p, q):
        (x1, y1), (x2, y2) = p, q
        t = D * x1 * x2 * y1 * y2 % P
        return ((x1 * y2 + y1 * x2) * pow(1 + t, P - 2, P) % P, (y1 * y2 + x1 * x2) * pow(1 - t, P - 2, P) % P)
    by = 4 * pow(5, P - 2, P) % P
    bx = 15112221349535400772501151409588531511454012693041857206046113283949847762202
    pts, cur = [(0, 1), (0, P - 1)], (bx, by)
    for _ in range(70):
        pts.append(cur); cur = aff_add(cur, (bx, by)) if rng.random() < 0.5 else aff_add(cur, cur)

    def ext(p):
        z = rng.randrange(1, P); x, y = p
        return [x * z % P, y * z % P, z, x * y * z % P]

    def rows(ps):
        return np.array([[(c >> (32 * i)) & 0xFFFFFFFF for c in e for i in range(8)] for e in ps], dtype=np.uint32)

    def affine(row):
        X, Y, Z, T = (sum(int(row[8 * c + i]) << (32 * i) for i in range(8)) % P for c in range(4))
        zi = pow(Z, P - 2, P)
        assert Z and X * Y % P == T * Z % P
        return (X * zi % P, Y * zi % P)
    pa = [rng.choice(pts) for _ in range(333)]; pb = [rng.choice(pts) for _ in range(333)]
    pb[:8] = pa[:8]                                               # P + P through the addition formula
    Ea, Eb = rows([ext(p) for p in pa]), rows([ext(p) for p in pb])
    for row, p in zip(ctx.selftest_oct(3, Ea, Eb), pa):
        assert affine(row) == aff_add(p, p)
    for row, p, q in zip(ctx.selftest_oct(4, Ea, Eb), pa, pb):
        assert affine(row) == aff_add(p, q)
