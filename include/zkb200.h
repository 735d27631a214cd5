/* zkb200.h - C ABI of libzkb200.so: B200-native (sm_100a) NTT/LDE/FRI-commit and MSM kernels that
 * sit behind the call sites crypto3-zk uses for its two data-parallel proving hot paths.
 *
 * The reference (NilFoundation/crypto3-zk, header-only C++) has no FFI/plugin layer: it reaches these
 * operations through C++ template entities of un-vendored sibling libraries.  Each entry point below
 * names the reference call site (file:line under /root/reference/include/nil/crypto3/) it serves and
 * the upstream entity it replaces; the C++ host templates in crypto3_zk_b200/host/ re-create those
 * entities on top of this ABI.
 *
 * Conventions
 *   - every function returns a zkb_status; nothing throws across the ABI;
 *   - field elements: canonical (non-Montgomery) integers < p as little-endian uint32 limbs,
 *     8 limbs (32 B) for the 254/255-bit fields, 12 limbs (48 B) for BLS12-381 Fq;
 *   - G1 points: affine (x || y), canonical limbs; the all-zero encoding is the point at infinity;
 *   - `mem` says where caller buffers live (ZKB_MEM_HOST / ZKB_MEM_DEVICE); host buffers are copied
 *     with cudaMemcpyAsync on the given stream (pin them for full PCIe rate);
 *   - `stream` is a cudaStream_t (NULL = legacy default stream).  Calls are stream-ordered;
 *     functions that return a result to host memory synchronise the stream before returning;
 *   - a context owns per-device caches (twiddle tables, scratch) and is not thread-safe: use one
 *     context per thread/GPU, like the reference's lazily cached evaluation_domain objects.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns ZKB_ERR_NO_DEVICE.
 */
#ifndef ZKB200_H
#define ZKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ZKB_OK = 0,
    ZKB_ERR_INVALID_ARGUMENT = 1, /* maps to std::invalid_argument in the host templates */
    ZKB_ERR_DOMAIN_TOO_LARGE = 2, /* 2^log_n exceeds the field's two-adicity */
    ZKB_ERR_CUDA = 3,             /* maps to std::runtime_error */
    ZKB_ERR_OUT_OF_MEMORY = 4,
    ZKB_ERR_NO_DEVICE = 5,
    ZKB_ERR_UNSUPPORTED = 6
} zkb_status;

/* arithmetic_params<F> of the fields on the hot path (ids shared with crypto3_zk_b200/fields.py) */
typedef enum {
    ZKB_FIELD_BLS12_381_FR = 0,
    ZKB_FIELD_BN254_FR = 1,
    ZKB_FIELD_PALLAS_FP = 2, /* Pallas base field (= Vesta scalar field) */
    ZKB_FIELD_PALLAS_FQ = 3, /* Pallas scalar field */
    ZKB_FIELD_BLS12_381_FQ = 4, /* coordinate field, no NTT */
    ZKB_FIELD_BN254_FQ = 5      /* coordinate field, no NTT */
} zkb_field;

/* G2 groups (coordinates in Fq2 = Fq[u]/(u^2+1), limbs c0 || c1): the B_query / knowledge-commitment side of
 * the Groth16 prover (r1cs_gg_ppzksnark/prover.hpp:113-119, knowledge_commitment_multiexp.hpp:107) and the
 * G2 commitment keys of ipp2 (ipp2/prover.hpp:209-212). */
typedef enum {
    ZKB_CURVE_BLS12_381_G1 = 0, ZKB_CURVE_BN254_G1 = 1, ZKB_CURVE_PALLAS = 2,
    ZKB_CURVE_BLS12_381_G2 = 3, ZKB_CURVE_BN254_G2 = 4
} zkb_curve;

typedef enum { ZKB_HASH_KECCAK_256 = 0, ZKB_HASH_SHA2_256 = 1, ZKB_HASH_KECCAK_512 = 2 } zkb_hash;

typedef enum { ZKB_MEM_HOST = 0, ZKB_MEM_DEVICE = 1 } zkb_mem;

typedef struct zkb_ctx zkb_ctx;
typedef struct zkb_msm_bases zkb_msm_bases;
typedef struct zkb_merkle_tree zkb_merkle_tree;
typedef struct zkb_sparse_matrix zkb_sparse_matrix;

/* ---- library / context ------------------------------------------------------------------- */
const char *zkb_version(void);
const char *zkb_status_string(int status);
int zkb_device_count(void);
int zkb_ctx_create(int device, zkb_ctx **out);
void zkb_ctx_destroy(zkb_ctx *ctx);
/* message of the most recent failing call on this context that recorded one ("" if none); wrappers read it right
 * after a non-zero status and then call zkb_ctx_clear_error so a later failure without a message cannot show it */
const char *zkb_ctx_last_error(const zkb_ctx *ctx);
void zkb_ctx_clear_error(zkb_ctx *ctx);
/* upper bound for internal scratch (bytes; default 6 GiB): batches are processed in chunks */
int zkb_ctx_set_scratch_limit(zkb_ctx *ctx, uint64_t bytes);
/* drops cached twiddle tables and scratch */
int zkb_ctx_release_caches(zkb_ctx *ctx);
/* number of kernels this context has launched so far (bench.py reports the delta) */
uint64_t zkb_ctx_kernel_launches(const zkb_ctx *ctx);

/* ---- field descriptors (host only, no device needed) --------------------------------------- */
int zkb_field_limbs(int field);        /* uint32 limbs per element, 0 if unknown */
int zkb_field_two_adicity(int field);
/* writes arithmetic_params<F>::multiplicative_generator (canonical limbs) */
int zkb_field_generator(int field, uint32_t *out);
/* math::unity_root<F>(2^log_n): the omega of make_evaluation_domain<F>(2^log_n) (canonical limbs).
 * Serves evaluation_domain::get_domain_element (basic_fri.hpp:783, fold_polynomial.hpp:83). */
int zkb_field_unity_root(int field, int log_n, uint32_t *out);

/* curve_type::g1_type<>::value_type::one(): the generator (affine x || y, canonical limbs); host only */
int zkb_curve_generator(int curve, uint32_t *out_affine);

/* ---- NTT: math::evaluation_domain<F>::fft / inverse_fft + math::multiply_by_coset ------------- */
/* Replaces basic_radix2_domain<F>::fft / inverse_fft as called from
 *   zk/snark/reductions/r1cs_to_qap.hpp:250,252,270,276,293,299,310
 *   zk/commitments/detail/polynomial/basic_fri.hpp:453 (through polynomial_dfs::resize)
 *   zk/snark/arithmetization/plonk/detail/column_polynomial.hpp:53
 * `batch` independent vectors of 2^log_n elements, contiguous ([batch][2^log_n]); natural order in and
 * out; out may equal in.  inverse != 0: inverse transform including the 1/n factor.
 * coset_shift (host pointer, canonical limbs, or NULL):
 *   forward: a[i] *= g^i first   == multiply_by_coset(a, g); fft(a)       (r1cs_to_qap.hpp:266-270)
 *   inverse: a[i] *= g^-i after  == inverse_fft(a); multiply_by_coset(a, g^-1) (r1cs_to_qap.hpp:310-315)
 */
int zkb_ntt(zkb_ctx *ctx, int field, int log_n, uint32_t batch, const void *in, void *out, int inverse,
            const uint32_t *coset_shift, int mem, void *stream);

/* ---- LDE: math::polynomial_dfs<V>::resize(2^log_n_out, nullptr, D) ---------------------------- */
/* = inverse_fft on the 2^log_n_in subgroup, zero-pad, fft on the 2^log_n_out subgroup
 * (basic_fri.hpp:369-371,451-455; gates_argument.hpp:120).  in: [batch][2^log_n_in],
 * out: [batch][2^log_n_out]; log_n_out >= log_n_in >= 1. */
int zkb_lde(zkb_ctx *ctx, int field, int log_n_in, int log_n_out, uint32_t batch, const void *in, void *out,
            int mem, void *stream);
/* the same resize on DEVICE buffers, also returning the coefficient form it passes through ([batch][2^log_n_in]):
 * lpc's eval_polys / proof_eval call `.coefficients()` on every committed polynomial again (batched_commitment.hpp:185,
 * lpc.hpp:142) - a scheme that keeps this by-product of commit() skips those inverse transforms */
int zkb_lde_with_coefficients(zkb_ctx *ctx, int field, int log_n_in, int log_n_out, uint32_t batch, const void *in_device,
                              void *out_device, void *coefficients_out_device, void *stream);

/* ---- pointwise helpers of the QAP witness map (r1cs_to_qap.hpp:256-262,280-285,301-321) -------- */
typedef enum {
    ZKB_VEC_MUL = 0,      /* out[i] = a[i] * b[i]                       (:283-285) */
    ZKB_VEC_SUB = 1,      /* out[i] = a[i] - b[i]                       (:304-306) */
    ZKB_VEC_ADD = 2,      /* out[i] = a[i] + b[i]                       (:319-321) */
    ZKB_VEC_MUL_SUB_SCALE = 3 /* out[i] = (a[i]*b[i] - c[i]) * s  (s = 1/Z(g): divide_by_z_on_coset, :283-308) */
} zkb_vec_op;
int zkb_vec(zkb_ctx *ctx, int field, int op, uint64_t n, const void *a, const void *b, const void *c,
            const uint32_t *scalar, void *out, int mem, void *stream);

/* ---- FRI fold: commitments::detail::fold_polynomial (dfs form) -------------------------------- */
/* zk/commitments/detail/polynomial/fold_polynomial.hpp:68-93:
 * out[i] = 1/2 * ((1 + alpha w^-i) f[i] + (1 - alpha w^-i) f[i + n/2]),  i < n/2, n = 2^log_n,
 * w = unity_root(2^log_n).  alpha: host pointer, canonical limbs. */
int zkb_fri_fold(zkb_ctx *ctx, int field, int log_n, const void *f, const uint32_t *alpha, void *out, int mem,
                 void *stream);

/* ---- LPC commit: zk::algorithms::precommit<FRI>(container<polynomial_dfs>, D, fri_step) -------- */
/* zk/commitments/detail/polynomial/basic_fri.hpp:445-496 + lpc.hpp:101-106.
 * polys: [batch][2^log_n_in] evaluations; every polynomial is resized to |D| = 2^log_n_out, leaves are
 * packed as in :466-492 (big-endian 32-byte elements, polynomial-major inside a leaf) and hashed into a
 * binary Merkle tree (containers::make_merkle_tree<Hash,2>).  The tree stays on the device behind
 * `tree_out` (may be NULL to discard); root_out (host, digest bytes) receives lpc::commit's result. */
int zkb_lpc_commit(zkb_ctx *ctx, int field, int hash, int log_n_in, int log_n_out, int fri_step, uint32_t batch,
                   const void *polys, int mem, uint8_t *root_out, zkb_merkle_tree **tree_out, void *stream);
/* Merkle tree only, over already extended evaluations [batch][2^log_n] (basic_fri.hpp:732 path) */
int zkb_merkle_commit(zkb_ctx *ctx, int field, int hash, int log_n, int fri_step, uint32_t batch,
                      const void *evals, int mem, uint8_t *root_out, zkb_merkle_tree **tree_out, void *stream);
/* root of the binary tree over `count` (a power of two) child digests given on the HOST: the top log2(count) levels
 * of a tree whose subtrees were committed separately (one per GPU, leaf ranges in rank order) */
int zkb_merkle_root_of_digests(zkb_ctx *ctx, int hash, uint32_t count, const uint8_t *digests, uint8_t *root_out,
                               void *stream);
int zkb_merkle_digest_bytes(int hash);
uint64_t zkb_merkle_leaves(const zkb_merkle_tree *tree);
/* authentication path of leaf `index`: `depth` sibling digests, leaf level first
 * (containers::merkle_proof<Hash,2>(tree, index), basic_fri.hpp:526-531) */
int zkb_merkle_path(zkb_ctx *ctx, const zkb_merkle_tree *tree, uint64_t index, uint8_t *path_out);
/* the paths of `count` leaves at once (the FRI query phase opens lambda leaves per tree, basic_fri.hpp:846-862):
 * one gather kernel and one copy; paths_out: host, count x depth digests */
int zkb_merkle_paths(zkb_ctx *ctx, const zkb_merkle_tree *tree, uint32_t count, const uint64_t *indices, uint8_t *paths_out,
                     void *stream);
void zkb_merkle_free(zkb_merkle_tree *tree);

/* ---- FRI commit phase: zk::algorithms::proof_eval<FRI>, "Commit phase" ---------------------------- */
/* zk/commitments/detail/polynomial/basic_fri.hpp:706-737.  f: the combined polynomial Q in evaluation form on
 * D[0] (2^log_n elements).  For every round i < rounds: tree_i = precommit<FRI>(f, D[t], step_list[i]); its root
 * goes to the caller's transcript, which answers with step_list[i] folding challenges; f is folded that many
 * times (fold_polynomial(f, alpha_t, D[t]), t running over all steps).  The transcript stays on the reference
 * side of the boundary: `challenge` is called once per round, on the calling thread, with the root digest, and
 * writes `count` = step_list[round] field elements (canonical limbs, 8 per element) to alphas_out; a non-zero
 * return aborts the call with ZKB_ERR_INVALID_ARGUMENT.
 *   roots_out      host, rounds digests (fri_roots)
 *   trees_out      NULL, or `rounds` handles (fri_trees; tree 0 is combined_Q_precommitment); free with zkb_merkle_free
 *   fs_device_out  NULL, or DEVICE memory that receives f after every round back to back (fs[1..rounds]):
 *                  sum_i 2^(log_n - step_list[0] - .. - step_list[i]) elements
 *   alphas_out     NULL, or host, sum(step_list) elements
 *   final_poly_out NULL, or host, 2^(log_n - sum(step_list)) elements: final_polynomial = f.coefficients() */
typedef int (*zkb_fri_challenge_fn)(void *user, uint32_t round, const uint8_t *root, uint32_t root_bytes, uint32_t count,
                                    uint32_t *alphas_out);
int zkb_fri_commit_phase(zkb_ctx *ctx, int field, int hash, int log_n, const void *f, int mem, const uint32_t *step_list,
                         uint32_t rounds, zkb_fri_challenge_fn challenge, void *user, uint8_t *roots_out,
                         zkb_merkle_tree **trees_out, void *fs_device_out, uint32_t *alphas_out, uint32_t *final_poly_out,
                         void *stream);

/* ---- grinding: proof_of_work<TranscriptHash, std::uint32_t>::generate ------------------------------------- */
/* zk/commitments/detail/polynomial/proof_of_work.hpp:47-68 over fiat_shamir_heuristic_sequential
 * (zk/transcript/fiat_shamir.hpp:152-164 absorb, :190-199 int_challenge): the smallest nonce >= start such that
 * (low32(H(H(state || be32(nonce)))) & mask) == 0.  `state` is the transcript's current digest (host,
 * zkb_merkle_digest_bytes(hash) bytes).  The reference starts from std::rand() and walks upwards one hash pair at a
 * time; any nonce that passes verifies (proof_of_work.hpp:70-78), the caller then absorbs be32(nonce) and draws the
 * int_challenge on its own transcript exactly as :64-66 does.  ZKB_ERR_UNSUPPORTED if no nonce in [start, 2^32) passes. */
int zkb_pow_grind(zkb_ctx *ctx, int hash, const uint8_t *state, uint32_t start, uint32_t mask, uint32_t *nonce_out,
                  void *stream);

/* ---- opening side of the LPC scheme: eval_polys and the combined quotient Q ------------------------- */
typedef enum { ZKB_POLY_COEFFICIENTS = 0, ZKB_POLY_DFS = 1 } zkb_poly_form;
/* polys_evaluator::eval_polys (zk/commitments/batched_commitment.hpp:176-190): _z(k,i,j) = polys[i].evaluate(point_j).
 * polys: [batch][n], coefficient form (any n) or evaluations on the 2^k subgroup (ZKB_POLY_DFS, n = 2^k: the
 * coefficients are recovered by one batched inverse NTT, as polynomial_dfs::evaluate does);
 * points: host, npoints elements; out: host, [batch][npoints] elements. */
int zkb_poly_evaluate(zkb_ctx *ctx, int field, int form, uint64_t n, uint32_t batch, const void *polys, int mem,
                      uint32_t npoints, const uint32_t *points, uint32_t *out, void *stream);
/* The FRI query phase opens every polynomial at the pairs (s, -s) of lambda query cosets
 * (basic_fri.hpp:819-834: `g_coeffs[k][polynomial_index].evaluate(s0)` / `.evaluate(s1)`): values of the coefficient-form
 * polynomials (device) at z and at -z for every point with ONE pass over the coefficients per four points.
 * out: host, [batch][npoints][2] elements (value at z, value at -z). */
int zkb_poly_evaluate_pm(zkb_ctx *ctx, int field, uint64_t n, uint32_t batch, const void *polys_device, uint32_t npoints,
                         const uint32_t *points, uint32_t *out, void *stream);
/* numerator of one point of lpc::proof_eval's combined Q (zk/commitments/polynomial/lpc.hpp:139-153, 163-176):
 * out[i] (+)= sum_j scalars[j] * polys[j][i] - [i == 0] * constant, with scalars[j] = theta^k for the polynomials
 * opened at the point and 0 for the others (skipped), constant = sum_j theta^k z_j (NULL = 0).
 * polys / out: DEVICE, coefficient form, [batch][n] / [n]; scalars: host [batch]; accumulate != 0 adds to out. */
int zkb_poly_lincomb(zkb_ctx *ctx, int field, uint64_t n, uint32_t batch, const void *polys_device, const uint32_t *scalars,
                     const uint32_t *constant, void *out_device, int accumulate, void *stream);
/* Q_normal / V with V = X - point (lpc.hpp:154,177): out = quotient (n-1 coefficients, out[n-1] = 0), the remainder
 * in(point) is dropped like upstream polynomial division does and returned through remainder_out (host, may be
 * NULL) so that callers can assert it is zero.  in / out: DEVICE, distinct buffers of n elements. */
int zkb_poly_div_linear(zkb_ctx *ctx, int field, uint64_t n, const void *in_device, const uint32_t *point, void *out_device,
                        uint32_t *remainder_out, void *stream);

/* ---- R1CS rows: cs.constraints[i].a/b/c.evaluate(full_variable_assignment) --------------------------- */
/* zk/snark/reductions/r1cs_to_qap.hpp:245-248, 289-291.  One CSR matrix per side of the constraint system (part of
 * the proving key, so it is uploaded once): row i = linear combination i, column j = variable j with column 0 the
 * constant 1 (x[0] = 1, x[1..] = primary || auxiliary input).  values: host, canonical limbs, nnz elements. */
int zkb_sparse_matrix_create(zkb_ctx *ctx, int field, uint64_t rows, uint64_t cols, const uint64_t *row_ptr,
                             const uint32_t *col_idx, const uint32_t *values, void *stream, zkb_sparse_matrix **out);
void zkb_sparse_matrix_free(zkb_sparse_matrix *m);
/* y[i] = sum_k values[k] * x[col_idx[k]] over row i; x: `cols` elements (host or device), y: DEVICE, `rows` elements */
int zkb_sparse_matvec(zkb_ctx *ctx, const zkb_sparse_matrix *m, const void *x, int x_mem, void *y_device, void *stream);

/* ---- MSM: algebra::multiexp<Method> / multiexp_with_mixed_addition<Method> --------------------- */
/* Call sites: zk/commitments/polynomial/kzg.hpp:146,414,433 ; kzg_v2.hpp:215 ;
 * zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/prover.hpp:108-139 ;
 * zk/commitments/polynomial/knowledge_commitment_multiexp.hpp:107.
 * Bases are uploaded once (an SRS / proving-key query vector is long-lived) ... */
int zkb_msm_bases_create(zkb_ctx *ctx, int curve, uint64_t n, const void *points_affine, int mem, void *stream,
                         zkb_msm_bases **out);
void zkb_msm_bases_free(zkb_msm_bases *bases);
uint64_t zkb_msm_bases_size(const zkb_msm_bases *bases);
/* Optional, once per base vector: builds the window table 2^(c w) * P_i (w < ceil((bits+1)/c)) next to the
 * bases so that later zkb_msm / zkb_msm_partial calls on them add every digit window into one bucket set.
 * Same results, fewer windows; costs ceil((bits+1)/c) times the memory of the bases and one inversion per
 * point and window.  window_bits = 0 picks c ~ log2 n (+2 below 2^17 points, +1 below 2^19; within [8, 22]); the
 * ceil((bits+1)/c) windows are then made equally wide.  Fails with ZKB_ERR_OUT_OF_MEMORY
 * (bases stay usable) when the table would exceed max_bytes.  For a commitment key that serves many commits
 * (kzg.hpp:100-118 builds it once per params_type). */
int zkb_msm_bases_precompute(zkb_ctx *ctx, zkb_msm_bases *bases, int window_bits, uint64_t max_bytes, void *stream);
/* result = sum_{i<n} scalars[i] * bases[offset+i]; scalars canonical limbs of the curve's scalar field;
 * result_affine: host buffer, (x||y) canonical limbs, all-zero for infinity.  Zero scalars are skipped and
 * unit scalars cost one mixed addition, so multiexp and multiexp_with_mixed_addition map to the same call. */
int zkb_msm(zkb_ctx *ctx, const zkb_msm_bases *bases, uint64_t offset, uint64_t n, const void *scalars, int mem,
            uint32_t *result_affine, void *stream);
/* same, but returns the result un-normalised: XYZZ, Montgomery form, 4 coordinates, written to HOST memory (the call
 * synchronises the stream: the window sums come back for the host-side window combine); zkb_msm_combine adds `count`
 * such partial results (e.g. one per GPU) on the host into an affine point.  Scalars must be canonical elements of
 * the scalar field (< 2^bits): a scalar with higher bits set fails the call with ZKB_ERR_INVALID_ARGUMENT. */
int zkb_msm_partial(zkb_ctx *ctx, const zkb_msm_bases *bases, uint64_t offset, uint64_t n, const void *scalars,
                    int mem, uint32_t *partial_xyzz_host, void *stream);
int zkb_msm_combine(int curve, uint32_t count, const uint32_t *partials_xyzz, uint32_t *result_affine);
/* what zkb_msm would do for n scalars on these bases: widest digit window in bits (2^(bits-1) buckets per set), number
 * of digit windows, number of bucket sets (1 with a window table).  Host only; for work models (bench.py). */
int zkb_msm_window_plan(const zkb_msm_bases *bases, uint64_t n, int *window_bits, int *windows, int *bucket_sets);
/* one-shot convenience (uploads bases every call) */
int zkb_msm_g1(zkb_ctx *ctx, int curve, uint64_t n, const void *points_affine, const void *scalars, int mem,
               uint32_t *result_affine, void *stream);

/* ---- multi-GPU (SURVEY.md 8(b) "zkb_*_multi taking a device list", 8(e)) ---------------------------- */
/* One host process drives several GPUs of a node: a zkb_multi owns one context per listed device (and enables peer
 * access between them).  Calls fan out on one host thread per device and return when all devices are done.
 *   MSM (kzg.hpp:146, prover.hpp:108-139): the bases are split into contiguous point ranges, one per device, resident
 *     there (zkb_msm_bases_multi_precompute adds the per-device window tables); zkb_msm_multi sends every device its
 *     slice of the host scalars, the XYZZ partial sums are added on the host.  scalars: host, n x 8 limbs, n <= bases.
 *   LPC commit (basic_fri.hpp:445-496): polynomials split over the devices for the resize; the extended evaluations are
 *     regrouped by leaf range with device-to-device copies, every device commits the subtree of its leaves, the top
 *     log2(devices) levels come from the subtree roots.  The device count must be a power of two dividing `batch` and
 *     the leaf count; polys: HOST, [batch][2^log_n_in].  Same root as zkb_lpc_commit on one device. */
typedef struct zkb_multi zkb_multi;
typedef struct zkb_msm_bases_multi zkb_msm_bases_multi;
int zkb_multi_create(const int *devices, uint32_t count, zkb_multi **out);
void zkb_multi_destroy(zkb_multi *m);
uint32_t zkb_multi_size(const zkb_multi *m);
zkb_ctx *zkb_multi_ctx(zkb_multi *m, uint32_t i);          /* the context of the i-th listed device (owned by m) */
const char *zkb_multi_last_error(const zkb_multi *m);
int zkb_msm_bases_multi_create(zkb_multi *m, int curve, uint64_t n, const void *points_affine_host, zkb_msm_bases_multi **out);
int zkb_msm_bases_multi_precompute(zkb_multi *m, zkb_msm_bases_multi *bases, int window_bits, uint64_t max_bytes_per_device);
void zkb_msm_bases_multi_free(zkb_msm_bases_multi *bases);
int zkb_msm_multi(zkb_multi *m, const zkb_msm_bases_multi *bases, uint64_t n, const void *scalars_host, uint32_t *result_affine);
int zkb_lpc_commit_multi(zkb_multi *m, int field, int hash, int log_n_in, int log_n_out, int fri_step, uint32_t batch,
                         const void *polys_host, uint8_t *root_out);

/* ---- synthetic inputs (benchmarks/tests; SURVEY.md 8(d) "generate on GPU") ------------------------ */
/* out_device[i] = table_a[i % m] + table_b[i / m]: n valid affine points from two small host tables
 * (m and ceil(n/m) affine points, canonical limbs).  out_device is DEVICE memory, n * 2 * coord limbs. */
int zkb_g1_grid_points(zkb_ctx *ctx, int curve, uint64_t n, uint32_t m, const void *table_a, const void *table_b,
                       void *out_device, void *stream);

/* ---- micro-benchmarks used for the integer-pipe roofline (bench.py) ------------------ */
/* runs `iters` dependent Montgomery multiplications per thread on `threads` threads; returns field-mul/s */
int zkb_bench_field_mul(zkb_ctx *ctx, int field, uint32_t blocks, uint32_t threads, uint32_t iters, double *muls_per_s);
/* bare IMAD.WIDE (mad.wide.u32) issue rate, 8 independent chains per thread, no field code: 32 * iters wide
 * multiply-adds per thread; returns wide multiply-adds/s.  The independent ceiling behind the product peaks. */
int zkb_bench_imad_wide(zkb_ctx *ctx, uint32_t blocks, uint32_t threads, uint32_t iters, double *wide_per_s);

/* ---- fixed-base batch exponentiation: the Groth16 generator ----------------------------------- */
/* algebra::batch_exp<G, Fr>(scalar_size, window, table, v) and the windowed_exp calls of kc_batch_exp
 * (zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/generator.hpp:167-225,
 *  zk/commitments/polynomial/knowledge_commitment_multiexp.hpp:110-205): out[i] = scalars[i] * base, affine; a zero scalar
 * gives the point at infinity (all-zero encoding - kc_batch_exp skips those entries, :129).  base_affine: host, (x || y)
 * canonical limbs; scalars / out: host or device by `mem`, n x 8 limbs / n points.  G1 and G2 curves. */
int zkb_batch_exp(zkb_ctx *ctx, int curve, uint64_t n, const uint32_t *base_affine, const void *scalars, void *out_affine, int mem,
                  void *stream);

/* ---- multiplicative scans: the grand product of the Placeholder permutation argument ---------- */
/* zk/snark/systems/plonk/placeholder/permutation_argument.hpp:104-133:
 *   g_v[i] = column_i + beta S_id[i] + gamma,  h_v[i] = column_i + beta S_sigma[i] + gamma   (i < ncols, all of n rows)
 *   V_P[0] = 1,  V_P[j] = V_P[j-1] * prod_i g_v[i][j-1] * (prod_i h_v[i][j-1]).inversed()
 * columns / s_id / s_sigma: DEVICE, [ncols][n] canonical elements (the columns selected by global_indices, in that
 * order); beta, gamma: host; v_out: DEVICE, n elements.  The reference inverts once per row; here one inversion serves
 * the whole column (prefix and suffix product scans).  ZKB_ERR_INVALID_ARGUMENT if a denominator is zero (the
 * reference's inversed() of zero is undefined). */
int zkb_permutation_grand_product(zkb_ctx *ctx, int field, uint64_t n, uint32_t ncols, const void *columns_device,
                                  const void *s_id_device, const void *s_sigma_device, const uint32_t *beta, const uint32_t *gamma,
                                  void *v_out_device, void *stream);
/* compute_V_L of the lookup argument (zk/snark/systems/plonk/placeholder/lookup_argument.hpp:375-409): V_L[0] = 1 and for
 * k = 1 .. usable_rows (< n)
 *   V_L[k] = V_L[k-1] (1+beta)^n_inputs prod_i (gamma + input_i[k-1]) prod_i (part1 + value_i[k-1] + beta value_i[k])
 *                     / prod_i (part1 + sorted_i[k-1] + beta sorted_i[k]),        part1 = (1 + beta) gamma,
 * V_L[k] = 0 for k > usable_rows.  inputs / values / sorted: DEVICE, [count][n] canonical elements (reduced_input,
 * reduced_value, sorted); v_out: DEVICE, n elements. */
int zkb_lookup_grand_product(zkb_ctx *ctx, int field, uint64_t n, uint64_t usable_rows, uint32_t n_inputs, const void *inputs_device,
                             uint32_t n_values, const void *values_device, uint32_t n_sorted, const void *sorted_device,
                             const uint32_t *beta, const uint32_t *gamma, void *v_out_device, void *stream);
/* its two building blocks on DEVICE vectors of n canonical elements: out[i] = prod_{j < i} in[j] (exclusive != 0, out[0] = 1)
 * or prod_{j <= i} in[j]; out[i] = in[i]^-1 (ZKB_ERR_INVALID_ARGUMENT if some in[i] is zero).  The lookup argument's
 * V_L (lookup_argument.hpp) is the same pair of operations. */
int zkb_prefix_product(zkb_ctx *ctx, int field, uint64_t n, const void *in_device, void *out_device, int exclusive, void *stream);
int zkb_batch_inverse(zkb_ctx *ctx, int field, uint64_t n, const void *in_device, void *out_device, void *stream);

/* ---- Placeholder argument builders: expressions over an extended domain, quotient, lookup sort ------ */
/* A postfix program evaluated once per point i < n (n a power of two) over DEVICE column arrays [ncols][n] of canonical
 * elements: the gate expression of zk/snark/systems/plonk/placeholder/gates_argument.hpp:76-217 (selector * sum theta^k
 * constraint, times the mask) and the F parts of the permutation / lookup arguments (permutation_argument.hpp:170-215) are
 * such expressions.  The arrays hold the columns on ONE coset w_E^j H of the extended domain (one zkb_ntt with a coset
 * shift per coset, from the coefficient form); a rotation by r rows is the index shift (i + r) mod n.  The result goes to
 * out[out_offset + i * out_stride] (accumulate != 0 adds): offset j, stride D interleaves the D cosets into the evaluation
 * form on the size-(D n) subgroup that polynomial_dfs holds upstream.  At most 16 stack slots. */
typedef enum {
    ZKB_EXPR_PUSH_COL = 0,   /* a = column, b = rotation in rows (may be negative) */
    ZKB_EXPR_PUSH_CONST = 1, /* a = index into `constants` */
    ZKB_EXPR_ADD = 2, ZKB_EXPR_SUB = 3, ZKB_EXPR_MUL = 4, /* pop b, pop a, push a (op) b */
    ZKB_EXPR_NEG = 5
} zkb_expr_op;
typedef struct { uint32_t op; uint32_t a; int32_t b; } zkb_expr_instr;
int zkb_expr_eval(zkb_ctx *ctx, int field, uint64_t n, uint32_t ncols, const void *cols_device, const zkb_expr_instr *program,
                  uint32_t n_instr, const uint32_t *constants, uint32_t nconst, void *out_device, uint64_t out_stride,
                  uint64_t out_offset, int accumulate, void *stream);
/* quotient_polynomial + split (placeholder/prover.hpp:220-283, detail::split_polynomial :47-70): T = F / (X^n - 1) for F
 * in COEFFICIENT form (2^log_ext coefficients, DEVICE; the remainder is dropped like upstream's polynomial division), cut
 * into `nchunks` chunks of n = 2^log_n coefficients (missing ones are zero), every chunk taken to evaluation form on the
 * n-subgroup (from_coefficients, :255-257).  out: DEVICE, [nchunks][n]. */
int zkb_quotient_split(zkb_ctx *ctx, int field, int log_n, int log_ext, const void *f_coefficients_device, uint32_t nchunks,
                       void *out_dfs_device, void *stream);
/* sort_polynomials of the lookup argument (lookup_argument.hpp:565-633): inputs / values = reduced_input / reduced_value,
 * DEVICE [count][n]; only rows < usable_rows take part.  sorted: DEVICE, [n_inputs + n_values][n], receives every table
 * value as often as it occurs in the table and the inputs together, in table order (a single zero first, zero runs once),
 * usable_rows values per column, sorted[i][usable_rows] = sorted[i+1][0], zero elsewhere.  Counting is a device hash table
 * instead of the reference's unordered_map.  ZKB_ERR_INVALID_ARGUMENT when a lookup input is not in the table (the reference
 * asserts, :577) or the values do not fit (equal table values that are not adjacent are emitted once per run). */
int zkb_lookup_sort(zkb_ctx *ctx, int field, uint64_t n, uint64_t usable_rows, uint32_t n_inputs, const void *inputs_device,
                    uint32_t n_values, const void *values_device, void *sorted_device, void *stream);

/* ---- compressed points of the Groth16 wire format ---------------------------------------------- */
/* Batched curve_element_serializer<bls12<381>>::octets_to_g1_point / octets_to_g2_point: the readers of
 * zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/marshalling.hpp:97-198 call it once per element of a proving key's query
 * vectors (:656-738) and every call is a square root in Fq (G1) or Fq2 (G2).  octets: n encodings of 48 (G1) / 96 (G2)
 * bytes, stride_bytes apart (a knowledge-commitment vector interleaves G2 | G1: stride 144); out: n affine points,
 * canonical limbs, (0, 0) for the point at infinity - the layout zkb_msm_bases_create takes.  status_out (optional, n
 * bytes, same memory space): 0 ok, 1 infinity, 2 not in compressed form, 3 infinity flag with a payload, 4 coordinate not
 * reduced, 5 x is not the abscissa of a curve point.  Any status >= 2 makes the call return ZKB_ERR_INVALID_ARGUMENT
 * (the reader's status_type::invalid_msg_data); the message names the first such point.  curve: ZKB_CURVE_BLS12_381_G1 /
 * _G2 (ZKB_ERR_UNSUPPORTED otherwise: upstream serializes no other curve in this format). */
int zkb_points_decompress(zkb_ctx *ctx, int curve, uint64_t n, const uint8_t *octets, uint64_t stride_bytes,
                          void *points_affine_out, uint8_t *status_out, int mem, void *stream);

/* ---- device buffers for host templates ------------------------------------------------------------ */
/* lpc_commitment_scheme keeps its polynomials as members between commit / eval_polys / proof_eval
 * (zk/commitments/polynomial/lpc.hpp:66-200, batched_commitment.hpp:60-250: `_polys`, `_z`); a host template over this ABI
 * keeps them in device buffers.  zkb_buf_copy synchronises the stream whenever a host buffer is involved;
 * zkb_gather reads `count` 32-byte field elements src[indices[i]] into host memory with one kernel and one copy (the FRI
 * query phase takes 2 lambda values from every retained f_i, basic_fri.hpp:880-887). */
int zkb_buf_alloc(zkb_ctx *ctx, uint64_t bytes, void **device_out);
void zkb_buf_free(zkb_ctx *ctx, void *device_ptr);
int zkb_buf_copy(zkb_ctx *ctx, void *dst, int dst_mem, const void *src, int src_mem, uint64_t bytes, void *stream);
int zkb_buf_zero(zkb_ctx *ctx, void *device_ptr, uint64_t bytes, void *stream);
int zkb_gather(zkb_ctx *ctx, const void *src_device, uint32_t count, const uint64_t *indices, uint32_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
