// tests/hostemu/fe_emu.cpp -- TEST HARNESS ONLY (never part of the product library).
// Compiles the device arithmetic headers for the host (carry chains as portable C) and exposes them to pytest, so
// the logic above the asm chains is checked on the CPU against Python big integers / the oracle before GPU time.
#include "../../xelis_he_b200/csrc/ge25519.cuh"
#include "../../xelis_he_b200/csrc/sc25519.cuh"
#include <string.h>
using namespace xhe;
extern "C" {
void emu_fe_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  fe x, y, r; memcpy(x.v, a, 32); memcpy(y.v, b, 32);
  switch (op) {
    case 0: r = fe_add(x, y); break; case 1: r = fe_sub(x, y); break; case 2: r = fe_mul(x, y); break; case 3: r = fe_sq(x); break;
    case 4: r = fe_freeze(x); break; case 5: r = fe_invert(x); break; case 6: r = fe_pow22523(x); break; case 7: r = fe_neg(x); break;
    default: r = fe_zero();
  }
  memcpy(out, r.v, 32);
}
void emu_sc_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  sc x = sc_frombytes(a), y = sc_frombytes(b), r;
  switch (op) {
    case 0: r = sc_add(x, y); break; case 1: r = sc_sub(x, y); break; case 2: r = sc_mul(x, y); break;
    case 3: r = sc_from_mont(sc_mont_invert(sc_to_mont(x))); break; case 4: r = sc_neg(x); break;
    case 5: r = sc_reduce512(x, y); break; case 6: r = sc_reduce256(x); break; default: r = sc_zero();
  }
  sc_tobytes(out, r);
}
int emu_decode(const uint8_t* enc, uint8_t* xy) { ge_aff a; bool ok = ristretto_decode(a, enc); fe_tobytes(xy, a.x); fe_tobytes(xy + 32, a.y); return ok; }
// op 0: decode a, decode b, add, encode; 1: sub; 2: double a; 3: madd(a, niels(b)); 4: encode+affine roundtrip
int emu_point_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  ge_aff pa, pb; if (!ristretto_decode(pa, a) || !ristretto_decode(pb, b)) return 0;
  ge A = ge_from_affine(pa), B = ge_from_affine(pb), R;
  switch (op) { case 0: R = ge_add(A, B); break; case 1: R = ge_sub(A, B); break; case 2: R = ge_double(A); break; case 3: R = ge_madd(ge_double(A), niels_from_affine(pb)); break;
    default: { ge S = ge_add(ge_double(A), B); ge_aff n; uint8_t tmp[32]; ristretto_encode(tmp, S, &n); R = ge_from_affine(n); } }
  ristretto_encode(out, R); return 1;
}
void emu_from_uniform(const uint8_t* u, uint8_t* out) { ristretto_encode(out, ristretto_from_uniform(u)); }
}
