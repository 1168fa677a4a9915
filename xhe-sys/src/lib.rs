//! Raw bindings to `libxhe_cuda.so` (`include/xhe.h`): the drop-in boundary for the batch-verification hot path of
//! xelis-he.  Layouts mirror the C header field for field; `struct_size` must be `size_of::<T>()` (use `Default`), the
//! library refuses anything else with `XHE_E_ARG`.
//!
//! Call sites in the reference that bind here:
//! * `BatchCollector::verify` (`src/proofs.rs:49-67`)                 -> [`xhe_msm_vartime`]
//! * `Compressed*::decompress` in bulk (`src/compressed.rs:28-106`)   -> [`xhe_ristretto_decompress`]
//! * `apply_without_verify` balance algebra (`src/tx/verify.rs:545-619`) -> [`xhe_ct_update`]
//! * `Signature::verify` group part (`src/elgamal.rs:38-42`)          -> [`xhe_sig_r`]
//! * `Transaction::verify_batch` device part (`src/tx/verify.rs:487-517`) -> [`xhe_verify_batch`]
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const XHE_OK: i32 = 0;
pub const XHE_ERR_SIGNATURE: i32 = 1;
pub const XHE_ERR_DECOMPRESSION: i32 = 2;
pub const XHE_ERR_COMMITMENT_EQ_PROOF: i32 = 3;
pub const XHE_ERR_CT_VALIDITY_PROOF: i32 = 4;
pub const XHE_ERR_GENERIC_PROOF: i32 = 5;
pub const XHE_ERR_RANGE_PROOF: i32 = 6;
pub const XHE_ERR_TRANSCRIPT: i32 = 7;
pub const XHE_ERR_FORMAT: i32 = 8;
pub const XHE_ERR_INVALID_NONCE: i32 = 9;
pub const XHE_ERR_STATE: i32 = 10;
pub const XHE_ERR_PARSE: i32 = 11;
pub const XHE_E_ARG: i32 = -1;
pub const XHE_E_CUDA: i32 = -2;
pub const XHE_E_NOMEM: i32 = -3;
pub const XHE_E_NCCL: i32 = -4;
pub const XHE_E_CAPACITY: i32 = -5;
pub const XHE_OP_PLUS_AMOUNT: i64 = 1 << 50;
pub const XHE_OP_FROM_LEDGER: i64 = 1 << 49;

#[repr(C)]
pub struct xhe_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct xhe_ledger {
    _private: [u8; 0],
}
#[repr(C)]
pub struct xhe_ecdlp {
    _private: [u8; 0],
}

/// `xhe_batch` of include/xhe.h.  All pointers are HOST memory; point index 0 must be the identity encoding.
#[repr(C)]
pub struct xhe_batch {
    pub struct_size: u32,
    pub n_tx: u32,
    pub n_points: u32,
    pub points: *const u8,
    pub n_sigs: u32,
    pub sig_s: *const u8,
    pub sig_e: *const u8,
    pub sig_pk: *const u32,
    pub n_ops: u32,
    pub op_prev: *const i64,
    pub op_term_off: *const u32,
    pub op_terms: *const u32,
    pub op_amount: *const u64,
    pub max_chain: u32,
    pub n_eq: u32,
    pub eq_points: *const u32,
    pub eq_scalars: *const u8,
    pub n_val: u32,
    pub val_points: *const u32,
    pub val_scalars: *const u8,
    pub n_rp: u32,
    pub rp_m: *const u32,
    pub rp_point_off: *const u32,
    pub rp_points: *const u32,
    pub rp_scalars: *const u8,
    pub rp_chal_off: *const u32,
    pub rp_challenges: *const u8,
    // optional: device-side Fiat-Shamir
    pub fs_blobs: *const u8,
    pub fs_blob_off: *const u64,
    pub fs_plan: *const u32,
    pub fs_seed: [u8; 32],
    pub fs_index_base: u64,
    // optional: device-side layout
    pub layout_on_device: u32,
    pub n_region_b: u32,
    pub region_b: *const u8,
    // optional: device-resident ledger read by XHE_OP_FROM_LEDGER ops
    pub ledger: *const xhe_ledger,
}

impl Default for xhe_batch {
    fn default() -> Self {
        // all-zero is the valid "nothing optional" state of the C struct
        let mut b: xhe_batch = unsafe { std::mem::zeroed() };
        b.struct_size = std::mem::size_of::<xhe_batch>() as u32;
        b
    }
}

/// `xhe_verdict` of include/xhe.h.  The caller allocates the arrays it wants filled (null = not wanted).
#[repr(C)]
pub struct xhe_verdict {
    pub struct_size: u32,
    pub sigma_is_identity: i32,
    pub range_is_identity: i32,
    pub sigma_enc: [u8; 32],
    pub range_enc: [u8; 32],
    pub sigma_ext: [u8; 128],
    pub range_ext: [u8; 128],
    pub point_ok: *mut u8,
    pub sig_r: *mut u8,
    pub op_out: *mut u8,
    pub sig_ok: *mut u8,
    pub device_flags: u32,
    pub tx_flags: *mut u8,
}

impl Default for xhe_verdict {
    fn default() -> Self {
        let mut v: xhe_verdict = unsafe { std::mem::zeroed() };
        v.struct_size = std::mem::size_of::<xhe_verdict>() as u32;
        v
    }
}

extern "C" {
    // context = the lazy_statics H, BP_GENS, PC_GENS (src/elgamal.rs:16-24, src/proofs.rs:19-22)
    pub fn xhe_ctx_create(device: c_int, party_capacity: u32, out: *mut *mut xhe_ctx) -> i32;
    pub fn xhe_ctx_destroy(ctx: *mut xhe_ctx);
    pub fn xhe_ctx_party_capacity(ctx: *const xhe_ctx) -> u32;
    pub fn xhe_last_error(ctx: *const xhe_ctx) -> *const c_char;
    pub fn xhe_ctx_set_stream(ctx: *mut xhe_ctx, cuda_stream: *mut c_void) -> i32;
    pub fn xhe_ctx_sync(ctx: *mut xhe_ctx) -> i32;
    pub fn xhe_ctx_launch_count(ctx: *const xhe_ctx) -> u64;
    // host-buffer entry points
    pub fn xhe_ristretto_decompress(ctx: *mut xhe_ctx, enc: *const u8, n: usize, xy: *mut u8, ok: *mut u8) -> i32;
    pub fn xhe_ristretto_compress(ctx: *mut xhe_ctx, xy: *const u8, n: usize, enc: *mut u8) -> i32;
    pub fn xhe_ristretto_from_uniform(ctx: *mut xhe_ctx, uniform64: *const u8, n: usize, enc: *mut u8) -> i32;
    pub fn xhe_msm_vartime(ctx: *mut xhe_ctx, scalars: *const u8, enc_points: *const u8, n: usize, out_enc: *mut u8, is_identity: *mut i32) -> i32;
    pub fn xhe_ct_update(ctx: *mut xhe_ctx, bal: *const u8, delta: *const u8, sub: *const u8, n: usize, out: *mut u8, ok: *mut u8) -> i32;
    pub fn xhe_sig_r(ctx: *mut xhe_ctx, s: *const u8, e: *const u8, pk_enc: *const u8, n: usize, r_enc: *mut u8, ok: *mut u8) -> i32;
    // device-pointer entry points
    pub fn xhe_decompress_dev(ctx: *mut xhe_ctx, d_enc: *const c_void, n: usize, d_affine: *mut c_void, d_niels: *mut c_void, d_ok: *mut c_void) -> i32;
    pub fn xhe_compress_dev(ctx: *mut xhe_ctx, d_ext: *const c_void, n: usize, d_enc: *mut c_void) -> i32;
    pub fn xhe_ct_update_dev(ctx: *mut xhe_ctx, d_bal: *const c_void, d_delta: *const c_void, d_sub: *const c_void, n: usize, d_out: *mut c_void, d_ok: *mut c_void) -> i32;
    pub fn xhe_ct_update_resident_dev(ctx: *mut xhe_ctx, d_bal_ext: *mut c_void, d_delta_niels: *const c_void, d_sub: *const c_void, n: usize) -> i32;
    pub fn xhe_msm_workspace_bytes(ctx: *const xhe_ctx, n: usize) -> usize;
    pub fn xhe_msm_dev(ctx: *mut xhe_ctx, d_scalars: *const c_void, d_niels: *const c_void, n: usize, d_workspace: *mut c_void, workspace_bytes: usize, d_out_enc32: *mut c_void, d_is_identity_u32: *mut c_void) -> i32;
    // Transaction::verify_batch, device part
    pub fn xhe_verify_batch(ctx: *mut xhe_ctx, batch: *const xhe_batch, verdict: *mut xhe_verdict) -> i32;
    pub fn xhe_batch_prepare(ctx: *mut xhe_ctx, batch: *const xhe_batch) -> i32;
    pub fn xhe_batch_run(ctx: *mut xhe_ctx) -> i32;
    pub fn xhe_batch_fetch(ctx: *mut xhe_ctx, verdict: *mut xhe_verdict) -> i32;
    // sharded batches
    pub fn xhe_combine_partials(ctx: *mut xhe_ctx, ext: *const u8, n: usize, out_enc: *mut u8, is_identity: *mut i32) -> i32;
    pub fn xhe_sum_encodings(ctx: *mut xhe_ctx, enc: *const u8, n: usize, out_enc: *mut u8, is_identity: *mut i32, all_valid: *mut i32) -> i32;
    // decoding of decrypted amounts: ECDLPInstance::decode / ElGamalSecretKey::decrypt (src/elgamal.rs:67-92, 140-145)
    pub fn xhe_ecdlp_create(ctx: *mut xhe_ctx, l1_bits: u32, out: *mut *mut xhe_ecdlp) -> i32;
    pub fn xhe_ecdlp_destroy(table: *mut xhe_ecdlp);
    pub fn xhe_ecdlp_table_bytes(table: *const xhe_ecdlp) -> usize;
    pub fn xhe_ecdlp_decode(table: *mut xhe_ecdlp, points: *const u8, n: usize, range_bits: u32, out_value: *mut i64, status: *mut u8) -> i32;
    pub fn xhe_decrypt_decode(table: *mut xhe_ecdlp, secret_key: *const u8, ciphertexts: *const u8, n: usize, range_bits: u32, out_value: *mut i64, status: *mut u8) -> i32;
    // device-resident ledger: BlockchainVerificationState backend with balances decompressed on the device
    pub fn xhe_ledger_create(ctx: *mut xhe_ctx, capacity: usize, out: *mut *mut xhe_ledger) -> i32;
    pub fn xhe_ledger_destroy(ledger: *mut xhe_ledger);
    pub fn xhe_ledger_size(ledger: *const xhe_ledger) -> usize;
    pub fn xhe_ledger_load(ledger: *mut xhe_ledger, keys: *const u8, cts: *const u8, n: usize, ok: *mut u8) -> i32;
    pub fn xhe_ledger_update(ledger: *mut xhe_ledger, keys: *const u8, deltas: *const u8, sub: *const u8, n: usize, status: *mut u8) -> i32;
    pub fn xhe_ledger_update_dense_dev(ledger: *mut xhe_ledger, d_delta_niels_planar: *const c_void, d_sub: *const c_void) -> i32;
    pub fn xhe_ledger_export(ledger: *mut xhe_ledger, keys: *const u8, n: usize, out_cts: *mut u8, found: *mut u8) -> i32;
    pub fn xhe_ledger_commit_batch(ledger: *mut xhe_ledger, ctx: *mut xhe_ctx, slots: *const u32, ops: *const u32, n: usize) -> i32;
    pub fn xhe_ledger_slot(ledger: *const xhe_ledger, key64: *const u8) -> u32;
    pub fn xhe_ledger_snapshot(ledger: *mut xhe_ledger) -> i32;
    pub fn xhe_ledger_restore(ledger: *mut xhe_ledger) -> i32;
    pub fn xhe_ledger_device_table(ledger: *const xhe_ledger, plane_stride_points: *mut usize) -> *mut c_void;
}

#[cfg(test)]
mod tests {
    use super::*;
    // layout pins against include/xhe.h on x86-64 / aarch64 (LP64): a drift shows up here before it shows up as XHE_E_ARG
    #[test]
    fn struct_sizes_match_the_header() {
        assert_eq!(std::mem::size_of::<xhe_batch>(), 288);
        assert_eq!(std::mem::size_of::<xhe_verdict>(), 384);
    }
}
