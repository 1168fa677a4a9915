"""GPU parity for the Pippenger MSM (config 2) through the C ABI: byte-identical encodings against the CPU oracle at
sizes it can do, an exact known-answer construction at full size, and size-independent algebraic properties."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
L = 2**252 + 27742317777372353535851937790883648493
G = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")


@pytest.fixture(scope="module")
def ctx():
    import xelis_he_b200 as xhe
    c = xhe.Ctx(0, party_capacity=2)
    yield c
    c.close()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 189, 190, 700, 4096, 65536])
def test_msm_matches_oracle(ctx, n):
    import oracle
    s, p = oracle.gen_msm_inputs(1000 + n, n, threads=8) if n else (b"", b"")
    want = oracle.msm(s, p) if n else bytes(32)
    got, ident = ctx.msm(s, p)
    assert got == want
    assert ident == (want == bytes(32))


def test_msm_edge_scalars_and_points(ctx):
    import oracle
    s, p = oracle.gen_msm_inputs(5, 64)
    sc = [s[32 * i:32 * i + 32] for i in range(64)]
    pt = [p[32 * i:32 * i + 32] for i in range(64)]
    # zero / one / l-1 scalars, identity points, duplicated points, P and -P cancelling
    sc[0] = bytes(32); sc[1] = (1).to_bytes(32, "little"); sc[2] = (L - 1).to_bytes(32, "little"); sc[3] = (2**252).to_bytes(32, "little")
    pt[4] = bytes(32); pt[5] = pt[6]; pt[7] = pt[6]
    S, P = b"".join(sc), b"".join(pt)
    assert ctx.msm(S, P)[0] == oracle.msm(S, P)
    # sum = 0: s*P + (l-s)*P
    neg = ((L - int.from_bytes(sc[10], "little")) % L).to_bytes(32, "little")
    enc, ident = ctx.msm(sc[10] + neg, pt[10] + pt[10])
    assert enc == bytes(32) and ident
    # the SURVEY appendix-B identity trap: G + (-G) is a Ristretto identity that is not (0:1:1:0) after Edwards addition
    enc, ident = ctx.msm((1).to_bytes(32, "little") * 2, G + bytes.fromhex("eaffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f"))
    assert enc == bytes(32) and ident


def test_msm_rejects_bad_arguments(ctx):
    import oracle
    import xelis_he_b200 as xhe
    s, p = oracle.gen_msm_inputs(6, 8)
    with pytest.raises(xhe.XheError):
        ctx.msm(s[:32 * 7] + L.to_bytes(32, "little"), p)                   # non-canonical scalar
    with pytest.raises(xhe.XheError):
        ctx.msm(s, p[:32 * 7] + (1).to_bytes(32, "little"))                 # invalid point encoding


@pytest.mark.parametrize("logn", [18, 20, 22])      # 2^22 = the largest size of BASELINE config 2
def test_msm_known_answer_full_size(ctx, logn):
    """sum s_i B_{j(i)} with B_j = b_j G (256 known multiples): expected = (sum s_i b_j(i)) G, exact at any n."""
    import oracle
    n = 1 << logn
    base_sc = [oracle.sc_reduce_wide(hashlib.shake_256(b"base%d" % j).digest(64)) for j in range(256)]
    bases = [oracle.scalarmult(b, G) for b in base_sc]
    rng = np.random.default_rng(logn)
    idx = rng.integers(0, 256, size=n, dtype=np.uint32)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    sc[:, 31] &= 0x0F                                                       # < 2^252 < l: canonical
    scalars = sc.tobytes()
    points = np.frombuffer(b"".join(bases), dtype=np.uint8).reshape(256, 32)[idx].tobytes()
    want = oracle.msm_expected_known_bases(scalars, idx, b"".join(base_sc))
    got, _ = ctx.msm(scalars, points)
    assert got == want


def test_msm_linearity_and_split(ctx):
    """size-independent properties at 2^17: MSM(s,P)+MSM(t,P) = MSM(s+t,P); halves sum to the whole."""
    import oracle
    n = 1 << 17
    rng = np.random.default_rng(3)
    uni = rng.integers(0, 256, size=(n, 64), dtype=np.uint8).tobytes()
    pts = ctx.from_uniform(uni)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 31] &= 0x07
    b = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); b[:, 31] &= 0x07
    sa, sb = a.tobytes(), b.tobytes()
    # a + b < 2^252 so plain integer addition of the byte strings is the scalar sum
    ab = (np.frombuffer(sa, dtype="<u8").reshape(n, 4).astype(object), np.frombuffer(sb, dtype="<u8").reshape(n, 4).astype(object))
    ssum = b"".join(((int.from_bytes(sa[32 * i:32 * i + 32], "little") + int.from_bytes(sb[32 * i:32 * i + 32], "little")) % L).to_bytes(32, "little") for i in range(n))
    ra, rb, rs = ctx.msm(sa, pts)[0], ctx.msm(sb, pts)[0], ctx.msm(ssum, pts)[0]
    assert oracle.point_add(ra, rb) == rs
    h = n // 2
    lo, hi = ctx.msm(sa[:32 * h], pts[:32 * h])[0], ctx.msm(sa[32 * h:], pts[32 * h:])[0]
    assert oracle.point_add(lo, hi) == ra
