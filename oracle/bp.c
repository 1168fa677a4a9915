/* oracle/bp.c -- TEST INFRASTRUCTURE.  See bp.h. */
#include "bp.h"
#include <stdlib.h>
#include <string.h>
static ge *gensG[XO_BP_PARTY_CAP], *gensH[XO_BP_PARTY_CAP];
static ge *gen_chain(char which, uint32_t party) { /* GeneratorsChain: SHAKE256("GeneratorsChain" || label), 64-byte blocks */
  ge *out = malloc(sizeof(ge) * XO_BP_N); xo_sponge sp; xo_sponge_init(&sp, 136); xo_sponge_absorb(&sp, "GeneratorsChain", 15);
  uint8_t label[5] = { (uint8_t)which, (uint8_t)party, (uint8_t)(party >> 8), (uint8_t)(party >> 16), (uint8_t)(party >> 24) };
  xo_sponge_absorb(&sp, label, 5); xo_sponge_finish(&sp, 0x1f);
  for (int i = 0; i < XO_BP_N; i++) { uint8_t u[64]; xo_sponge_squeeze(&sp, u, 64); ristretto_from_uniform(&out[i], u); }
  return out;
}
void xo_bp_ensure(int m) { for (int j = 0; j < m && j < XO_BP_PARTY_CAP; j++) { if (!gensG[j]) gensG[j] = gen_chain('G', j); if (!gensH[j]) gensH[j] = gen_chain('H', j); } }
const ge *xo_bp_G(int party, int i) { if (!gensG[party]) xo_bp_ensure(party + 1); return &gensG[party][i]; }
const ge *xo_bp_H(int party, int i) { if (!gensH[party]) xo_bp_ensure(party + 1); return &gensH[party][i]; }
static int ilog2(unsigned x) { int l = 0; while ((1u << l) < x) l++; return l; }
size_t xo_rp_size(int m) { return 32 * (9 + 2 * (size_t)ilog2(64 * m)); }
static void append_u64(xo_transcript *t, const char *l, uint64_t v) { xo_transcript_append_u64(t, l, v); }
static void rp_domain_sep(xo_transcript *t, uint64_t n, uint64_t m) { xo_transcript_append(t, "dom-sep", "rangeproof v1", 13); append_u64(t, "n", n); append_u64(t, "m", m); }
static void ipp_domain_sep(xo_transcript *t, uint64_t n) { xo_transcript_append(t, "dom-sep", "ipp v1", 6); append_u64(t, "n", n); }
static void sc_inner(sc *r, const sc *a, const sc *b, int n) { sc acc, t; sc_0(&acc); for (int i = 0; i < n; i++) { sc_mul(&t, &a[i], &b[i]); sc_add(&acc, &acc, &t); } *r = acc; }

int xo_rp_prove(uint8_t *out, const uint64_t *values, const sc *blind, int m, xo_transcript *t, xo_rng *rng) {
  if (m < 1 || m > XO_BP_PARTY_CAP || (m & (m - 1))) return -1;
  const int n = XO_BP_N, N = n * m, lg = ilog2(N); xo_bp_ensure(m);
  const ge *G = xo_G(), *H = xo_H();
  rp_domain_sep(t, n, m);
  for (int j = 0; j < m; j++) { sc v; sc_from_u64(&v, values[j]); sc s2[2] = { v, blind[j] }; ge p2[2] = { *G, *H }; ge V; ge_msm_vartime(&V, s2, p2, 2); uint8_t e[32]; ristretto_encode(e, &V); xo_transcript_append(t, "V", e, 32); }
  sc *sL = malloc(sizeof(sc) * N), *sR = malloc(sizeof(sc) * N), *l0 = malloc(sizeof(sc) * N), *r0 = malloc(sizeof(sc) * N), *r1 = malloc(sizeof(sc) * N);
  sc *a_bl = malloc(sizeof(sc) * m), *s_bl = malloc(sizeof(sc) * m);
  ge A, S; ge_identity(&A);
  { /* bit commitments */
    sc sum_a, sum_s; sc_0(&sum_a); sc_0(&sum_s);
    for (int j = 0; j < m; j++) {
      xo_rng_scalar(rng, &a_bl[j]); xo_rng_scalar(rng, &s_bl[j]); sc_add(&sum_a, &sum_a, &a_bl[j]); sc_add(&sum_s, &sum_s, &s_bl[j]);
      for (int i = 0; i < n; i++) { if ((values[j] >> i) & 1) ge_add(&A, &A, xo_bp_G(j, i)); else ge_sub(&A, &A, xo_bp_H(j, i)); xo_rng_scalar(rng, &sL[j*n+i]); xo_rng_scalar(rng, &sR[j*n+i]); }
    }
    ge tmp; ge_scalarmult(&tmp, &sum_a, H); ge_add(&A, &A, &tmp);
    sc *ss = malloc(sizeof(sc) * (2*N + 1)); ge *pp = malloc(sizeof(ge) * (2*N + 1));
    for (int k = 0; k < N; k++) { ss[k] = sL[k]; pp[k] = *xo_bp_G(k / n, k % n); ss[N + k] = sR[k]; pp[N + k] = *xo_bp_H(k / n, k % n); }
    ss[2*N] = sum_s; pp[2*N] = *H; ge_msm_vartime(&S, ss, pp, 2*N + 1); free(ss); free(pp);
  }
  uint8_t *o = out; ristretto_encode(o, &A); ristretto_encode(o + 32, &S);
  xo_transcript_append(t, "A", o, 32); xo_transcript_append(t, "S", o + 32, 32);
  sc y, z, zz, x; xo_challenge_scalar(t, "y", &y); xo_challenge_scalar(t, "z", &z); sc_mul(&zz, &z, &z);
  /* polynomials l(x) = l0 + l1 x (l1 = sL), r(x) = r0 + r1 x */
  sc t0, t1, t2, one, exp_y, zzj; sc_1(&one); sc_0(&t0); sc_0(&t1); sc_0(&t2); sc_1(&exp_y); zzj = zz;
  for (int j = 0; j < m; j++) {
    sc exp_2; sc_1(&exp_2);
    for (int i = 0; i < n; i++) {
      int k = j*n + i; sc aL, aR, tmp; sc_from_u64(&aL, (values[j] >> i) & 1); sc_sub(&aR, &aL, &one);
      sc_sub(&l0[k], &aL, &z);
      sc_add(&tmp, &aR, &z); sc_mul(&tmp, &tmp, &exp_y); sc_mul(&r0[k], &zzj, &exp_2); sc_add(&r0[k], &r0[k], &tmp);
      sc_mul(&r1[k], &exp_y, &sR[k]);
      sc_mul(&exp_y, &exp_y, &y); sc_add(&exp_2, &exp_2, &exp_2);
    }
    sc_mul(&zzj, &zzj, &z);
  }
  { sc *ls = malloc(sizeof(sc) * N), *rs = malloc(sizeof(sc) * N), tt;
    sc_inner(&t0, l0, r0, N); sc_inner(&t2, sL, r1, N);
    for (int k = 0; k < N; k++) { sc_add(&ls[k], &l0[k], &sL[k]); sc_add(&rs[k], &r0[k], &r1[k]); }
    sc_inner(&tt, ls, rs, N); sc_sub(&t1, &tt, &t0); sc_sub(&t1, &t1, &t2); free(ls); free(rs); }
  sc t1_bl, t2_bl; xo_rng_scalar(rng, &t1_bl); xo_rng_scalar(rng, &t2_bl);
  { sc s2[2] = { t1, t1_bl }; ge p2[2] = { *G, *H }; ge T; ge_msm_vartime(&T, s2, p2, 2); ristretto_encode(o + 64, &T);
    s2[0] = t2; s2[1] = t2_bl; ge_msm_vartime(&T, s2, p2, 2); ristretto_encode(o + 96, &T); }
  xo_transcript_append(t, "T_1", o + 64, 32); xo_transcript_append(t, "T_2", o + 96, 32);
  xo_challenge_scalar(t, "x", &x);
  sc xx, t_x, t_x_bl, e_bl, tmp; sc_mul(&xx, &x, &x);
  sc_mul(&t_x, &t2, &xx); sc_mul(&tmp, &t1, &x); sc_add(&t_x, &t_x, &tmp); sc_add(&t_x, &t_x, &t0);
  sc_mul(&t_x_bl, &t2_bl, &xx); sc_mul(&tmp, &t1_bl, &x); sc_add(&t_x_bl, &t_x_bl, &tmp);
  zzj = zz; sc_0(&e_bl);
  for (int j = 0; j < m; j++) { sc_mul(&tmp, &zzj, &blind[j]); sc_add(&t_x_bl, &t_x_bl, &tmp); sc_mul(&zzj, &zzj, &z); sc_mul(&tmp, &s_bl[j], &x); sc_add(&tmp, &tmp, &a_bl[j]); sc_add(&e_bl, &e_bl, &tmp); }
  sc_tobytes(o + 128, &t_x); sc_tobytes(o + 160, &t_x_bl); sc_tobytes(o + 192, &e_bl);
  xo_transcript_append(t, "t_x", o + 128, 32); xo_transcript_append(t, "t_x_blinding", o + 160, 32); xo_transcript_append(t, "e_blinding", o + 192, 32);
  sc w; xo_challenge_scalar(t, "w", &w);
  /* inner-product argument over a = l(x), b = r(x), generators G_i, y^-i H_i, Q = w B (scalars tracked per original generator) */
  sc *a = l0, *b = r0; for (int k = 0; k < N; k++) { sc_mul(&tmp, &sL[k], &x); sc_add(&a[k], &l0[k], &tmp); sc_mul(&tmp, &r1[k], &x); sc_add(&b[k], &r0[k], &tmp); }
  sc *gc = sL, *hc = sR; sc yinv, e; sc_invert(&yinv, &y); sc_1(&e); for (int k = 0; k < N; k++) { sc_1(&gc[k]); hc[k] = e; sc_mul(&e, &e, &yinv); }
  ipp_domain_sep(t, N);
  ge Q; ge_scalarmult(&Q, &w, G);
  sc *ms = malloc(sizeof(sc) * (2*N + 1)); ge *mp = malloc(sizeof(ge) * (2*N + 1)); ge *cp = malloc(sizeof(ge) * (N + 1));
  for (int k = 0; k < N; k++) { mp[k] = *xo_bp_G(k / n, k % n); mp[N + k] = *xo_bp_H(k / n, k % n); } mp[2*N] = Q;
  uint8_t *lr = o + 224; int np = N;
  while (np > 1) {
    int h = np / 2; sc cL, cR; sc_inner(&cL, a, b + h, h); sc_inner(&cR, a + h, b, h);
    for (int pass = 0; pass < 2; pass++) { /* pass 0: L, pass 1: R */
      /* only the N + 1 non-zero terms are passed to the MSM (G_i or H_i per index, plus Q) */
      for (int i = 0; i < N; i++) {
        int k = i % np;
        if (pass == 0) { if (k >= h) { sc_mul(&ms[i], &a[k - h], &gc[i]); cp[i] = mp[i]; } else { sc_mul(&ms[i], &b[k + h], &hc[i]); cp[i] = mp[N + i]; } }
        else           { if (k < h) { sc_mul(&ms[i], &a[k + h], &gc[i]); cp[i] = mp[i]; } else { sc_mul(&ms[i], &b[k - h], &hc[i]); cp[i] = mp[N + i]; } }
      }
      ms[N] = pass == 0 ? cL : cR; cp[N] = mp[2*N]; ge P; ge_msm_vartime(&P, ms, cp, N + 1); ristretto_encode(lr, &P); xo_transcript_append(t, pass == 0 ? "L" : "R", lr, 32); lr += 32;
    }
    sc u, uinv; xo_challenge_scalar(t, "u", &u); sc_invert(&uinv, &u);
    for (int k = 0; k < h; k++) { sc t1_, t2_; sc_mul(&t1_, &a[k], &u); sc_mul(&t2_, &a[k + h], &uinv); sc_add(&a[k], &t1_, &t2_); sc_mul(&t1_, &b[k], &uinv); sc_mul(&t2_, &b[k + h], &u); sc_add(&b[k], &t1_, &t2_); }
    for (int i = 0; i < N; i++) { int lo = (i % np) < h; sc_mul(&gc[i], &gc[i], lo ? &uinv : &u); sc_mul(&hc[i], &hc[i], lo ? &u : &uinv); }
    np = h;
  }
  sc_tobytes(lr, &a[0]); sc_tobytes(lr + 32, &b[0]);
  free(ms); free(mp); free(cp); free(sL); free(sR); free(l0); free(r0); free(r1); free(a_bl); free(s_bl);
  (void)lg; return 0;
}

int xo_rp_verify_batch_ex(const xo_rp_item *items, size_t n_items, xo_rng *rng, uint8_t out_enc[32], int no_decision) {
  const int n = XO_BP_N; int m_max = 0;
  for (size_t q = 0; q < n_items; q++) { if (items[q].m > XO_BP_PARTY_CAP) return XO_ERR_RANGE_PROOF; if (items[q].m > m_max) m_max = items[q].m; }
  xo_bp_ensure(m_max);
  int Nmax = n * m_max; sc *gs = calloc(Nmax + 1, sizeof(sc)), *hs = calloc(Nmax + 1, sizeof(sc)); sc base_s, blind_s; sc_0(&base_s); sc_0(&blind_s);
  size_t cap = 64, nd = 0; sc *ds = malloc(cap * sizeof(sc)); ge *dp = malloc(cap * sizeof(ge)); int rc = XO_OK;
#define PUSH(sv, pv) do { if (nd == cap) { cap *= 2; ds = realloc(ds, cap * sizeof(sc)); dp = realloc(dp, cap * sizeof(ge)); } ds[nd] = (sv); dp[nd] = (pv); nd++; } while (0)
  for (size_t q = 0; q < n_items && rc == XO_OK; q++) {
    const xo_rp_item *it = &items[q]; const uint8_t *pr = it->proof; int m = it->m, N = n * m; xo_transcript *t = it->t;
    if (it->len % 32 || it->len < 9 * 32 || ((it->len / 32 - 9) & 1)) { rc = XO_ERR_RANGE_PROOF; break; }
    int lg = (int)((it->len / 32 - 9) / 2);
    if (lg >= 32 || m < 1 || N != (1 << lg)) { rc = XO_ERR_RANGE_PROOF; break; }
    sc t_x, t_x_bl, e_bl, a, b; const uint8_t *lr = pr + 224, *ab = lr + 64 * lg;
    if (!sc_frombytes_canonical(&t_x, pr + 128) || !sc_frombytes_canonical(&t_x_bl, pr + 160) || !sc_frombytes_canonical(&e_bl, pr + 192) ||
        !sc_frombytes_canonical(&a, ab) || !sc_frombytes_canonical(&b, ab + 32)) { rc = XO_ERR_RANGE_PROOF; break; }
    rp_domain_sep(t, n, m);
    for (int j = 0; j < m; j++) xo_transcript_append(t, "V", it->commit_enc + 32 * j, 32);
    if (!xo_validate_and_append_point(t, "A", pr) || !xo_validate_and_append_point(t, "S", pr + 32)) { rc = XO_ERR_RANGE_PROOF; break; }
    sc y, z, zz, x, w; xo_challenge_scalar(t, "y", &y); xo_challenge_scalar(t, "z", &z); sc_mul(&zz, &z, &z);
    if (!xo_validate_and_append_point(t, "T_1", pr + 64) || !xo_validate_and_append_point(t, "T_2", pr + 96)) { rc = XO_ERR_RANGE_PROOF; break; }
    xo_challenge_scalar(t, "x", &x);
    xo_transcript_append(t, "t_x", pr + 128, 32); xo_transcript_append(t, "t_x_blinding", pr + 160, 32); xo_transcript_append(t, "e_blinding", pr + 192, 32);
    xo_challenge_scalar(t, "w", &w);
    sc c, rho; xo_rng_scalar(rng, &c); xo_rng_scalar(rng, &rho);   /* c: intra-proof weight; rho: cross-proof batch factor */
    ipp_domain_sep(t, N);
    sc u[32], uinv[32], usq[32], uinvsq[32], allinv;
    for (int k = 0; k < lg; k++) { if (!xo_validate_and_append_point(t, "L", lr + 64 * k) || !xo_validate_and_append_point(t, "R", lr + 64 * k + 32)) { rc = XO_ERR_RANGE_PROOF; break; } xo_challenge_scalar(t, "u", &u[k]); uinv[k] = u[k]; }
    if (rc != XO_OK) break;
    sc_batch_invert(uinv, lg, &allinv);
    for (int k = 0; k < lg; k++) { sc_mul(&usq[k], &u[k], &u[k]); sc_mul(&uinvsq[k], &uinv[k], &uinv[k]); }
    sc *s = malloc(sizeof(sc) * N); s[0] = allinv;
    for (int i = 1; i < N; i++) { int lgi = 31 - __builtin_clz(i), k = 1 << lgi; sc_mul(&s[i], &s[i - k], &usq[lg - 1 - lgi]); }
    ge A, S, T1, T2, pt;
    if (!ristretto_decode(&A, pr) || !ristretto_decode(&S, pr + 32) || !ristretto_decode(&T1, pr + 64) || !ristretto_decode(&T2, pr + 96)) { rc = XO_ERR_RANGE_PROOF; free(s); break; }
    sc tmp, tmp2, cx; PUSH(rho, A); sc_mul(&tmp, &rho, &x); PUSH(tmp, S); sc_mul(&cx, &c, &x); sc_mul(&tmp, &cx, &rho); PUSH(tmp, T1); sc_mul(&tmp, &cx, &x); sc_mul(&tmp, &tmp, &rho); PUSH(tmp, T2);
    for (int k = 0; k < lg && rc == XO_OK; k++) { if (!ristretto_decode(&pt, lr + 64 * k)) { rc = XO_ERR_RANGE_PROOF; break; } sc_mul(&tmp, &usq[k], &rho); PUSH(tmp, pt); }
    for (int k = 0; k < lg && rc == XO_OK; k++) { if (!ristretto_decode(&pt, lr + 64 * k + 32)) { rc = XO_ERR_RANGE_PROOF; break; } sc_mul(&tmp, &uinvsq[k], &rho); PUSH(tmp, pt); }
    if (rc != XO_OK) { free(s); break; }
    /* B_blinding: -e_blinding - c t_x_blinding ; B: w (t_x - a b) + c (delta - t_x) */
    sc_mul(&tmp, &c, &t_x_bl); sc_add(&tmp, &tmp, &e_bl); sc_neg(&tmp, &tmp); sc_mul(&tmp, &tmp, &rho); sc_add(&blind_s, &blind_s, &tmp);
    sc sum_y, sum_z, sum_2, p, delta; sc_0(&sum_y); sc_1(&p); for (int i = 0; i < N; i++) { sc_add(&sum_y, &sum_y, &p); sc_mul(&p, &p, &y); }
    sc_0(&sum_z); sc_1(&p); for (int j = 0; j < m; j++) { sc_add(&sum_z, &sum_z, &p); sc_mul(&p, &p, &z); }
    sc_from_u64(&sum_2, 0xffffffffffffffffULL);
    sc_sub(&delta, &z, &zz); sc_mul(&delta, &delta, &sum_y); sc_mul(&tmp, &zz, &z); sc_mul(&tmp, &tmp, &sum_2); sc_mul(&tmp, &tmp, &sum_z); sc_sub(&delta, &delta, &tmp);
    sc_mul(&tmp, &a, &b); sc_sub(&tmp, &t_x, &tmp); sc_mul(&tmp, &tmp, &w); sc_sub(&tmp2, &delta, &t_x); sc_mul(&tmp2, &tmp2, &c); sc_add(&tmp, &tmp, &tmp2); sc_mul(&tmp, &tmp, &rho); sc_add(&base_s, &base_s, &tmp);
    sc yinv, exp_yinv, zzj; sc_invert(&yinv, &y); sc_1(&exp_yinv); zzj = zz;
    for (int j = 0; j < m; j++) {
      sc exp_2; sc_1(&exp_2);
      for (int i = 0; i < n; i++) {
        int k = j * n + i;
        sc_mul(&tmp, &a, &s[k]); sc_add(&tmp, &tmp, &z); sc_neg(&tmp, &tmp); sc_mul(&tmp, &tmp, &rho); sc_add(&gs[k], &gs[k], &tmp);       /* -z - a s_k */
        sc_mul(&tmp, &zzj, &exp_2); sc_mul(&tmp2, &b, &s[N - 1 - k]); sc_sub(&tmp, &tmp, &tmp2); sc_mul(&tmp, &tmp, &exp_yinv); sc_add(&tmp, &tmp, &z);
        sc_mul(&tmp, &tmp, &rho); sc_add(&hs[k], &hs[k], &tmp);
        sc_mul(&exp_yinv, &exp_yinv, &yinv); sc_add(&exp_2, &exp_2, &exp_2);
      }
      sc_mul(&tmp, &c, &zzj); sc_mul(&tmp, &tmp, &rho); PUSH(tmp, it->commit_pts[j]);   /* V_j: c z^2 z^j */
      sc_mul(&zzj, &zzj, &z);
    }
    free(s);
  }
  if (rc == XO_OK) {
    for (int k = 0; k < Nmax; k++) { PUSH(gs[k], *xo_bp_G(k / n, k % n)); } for (int k = 0; k < Nmax; k++) { PUSH(hs[k], *xo_bp_H(k / n, k % n)); }
    PUSH(base_s, *xo_G()); PUSH(blind_s, *xo_H());
    ge r; ge_msm_vartime(&r, ds, dp, nd); if (out_enc) ristretto_encode(out_enc, &r);
    if (!no_decision && !ge_ristretto_is_identity(&r)) rc = XO_ERR_RANGE_PROOF;
  }
  free(gs); free(hs); free(ds); free(dp); return rc;
}
int xo_rp_verify_batch(const xo_rp_item *items, size_t n_items, xo_rng *rng, uint8_t out_enc[32]) { return xo_rp_verify_batch_ex(items, n_items, rng, out_enc, 0); }
