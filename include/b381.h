/*
 * b381.h -- C ABI of the B200-native BLS12-381 proving backend (libb381_cuda.so).
 *
 * This is the drop-in boundary for the hot path of riusricardo/midnight-bls12-381-cuda:
 * every entry point below is what the reference's FFI for this path binds, with plain
 * pointers and sizes only.  Citations are file:line under the reference root.
 *
 *   - the functions ICICLE's dispatcher calls after `register_*`   (icicle_backend_api.cuh:118-193,
 *     registered at src/backend/icicle_curve_api.cu:660-665 and src/backend/icicle_field_api.cu:344-352)
 *     keep the argument order and meaning of those std::function signatures, minus the unused
 *     `const Device&`;
 *   - the flat `extern "C"` test API of the reference keeps its exact names and signatures
 *     (icicle_curve_api.cu:679-706, src/field/ntt_kernels.cu:1911-1942,
 *     icicle_field_api.cu:363-383, src/field/vec_ops.cu:393-524,693-838).
 *
 * Data layout (core/types.rs:89-108,148-270): little-endian u64 limbs, Montgomery form unless a
 * flag says otherwise; Fr 32 B, Fq 48 B, Fq2 = c0||c1, affine = x||y with infinity (0,0),
 * projective = X||Y||Z.  Pointers must be 16-byte aligned (every cudaMalloc / Rust Vec<Fq> /
 * DeviceVec allocation is).
 *
 * Error model: return code only (values of upstream icicle::eIcicleError,
 * include/icicle/errors.h:37-53); nothing throws across the boundary.
 * There is NO CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef B381_H
#define B381_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes = icicle::eIcicleError (include/icicle/errors.h:37-53) ---- */
enum {
  B381_SUCCESS = 0,
  B381_INVALID_DEVICE = 1,
  B381_OUT_OF_MEMORY = 2,
  B381_INVALID_POINTER = 3,
  B381_ALLOCATION_FAILED = 4,
  B381_DEALLOCATION_FAILED = 5,
  B381_COPY_FAILED = 6,
  B381_SYNCHRONIZATION_FAILED = 7,
  B381_STREAM_CREATION_FAILED = 8,
  B381_STREAM_DESTRUCTION_FAILED = 9,
  B381_API_NOT_IMPLEMENTED = 10,
  B381_INVALID_ARGUMENT = 11,
  B381_BACKEND_LOAD_FAILED = 12,
  B381_LICENSE_CHECK_ERROR = 13,
  B381_UNKNOWN_ERROR = 14
};

/* ---- element types (bls12-381/include/field.cuh:197-336, point.cuh:81-125,286-318,455-525) ---- */
typedef struct { uint64_t l[4]; } b381_fr;                 /* 32 B  */
typedef struct { uint64_t l[6]; } b381_fq;                 /* 48 B  */
typedef struct { b381_fq c0, c1; } b381_fq2;               /* 96 B  */
typedef struct { b381_fq x, y; } b381_g1_affine;           /* 96 B  */
typedef struct { b381_fq x, y, z; } b381_g1_projective;    /* 144 B */
typedef struct { b381_fq2 x, y; } b381_g2_affine;          /* 192 B */
typedef struct { b381_fq2 x, y, z; } b381_g2_projective;   /* 288 B */

/* ---- config structs: byte-compatible with the reference's (icicle_types.cuh:102-113,136-140,
 *      155-169,194-201).  The ICICLE glue in csrc/icicle/ reinterpret_casts to these. ---- */
typedef struct {
  void* stream;                    /* cudaStream_t */
  int precompute_factor;
  int c;                           /* window bits, 0 = auto */
  int bitsize;                     /* scalar bits; 0 = 255, the bit length of r (the reference reads 0 as 256 = 4 limbs,
                                      msm_kernels.cu:651: same digits for canonical scalars, one more empty window) */
  int batch_size;
  bool are_points_shared_in_batch;
  bool are_scalars_on_device;
  bool are_scalars_montgomery_form;
  bool are_points_on_device;
  bool are_points_montgomery_form;
  bool are_results_on_device;
  bool is_async;
  void* ext;
} b381_msm_config;

enum { B381_NTT_FORWARD = 0, B381_NTT_INVERSE = 1 };                       /* NTTDir   */
enum { B381_kNN = 0, B381_kNR = 1, B381_kRN = 2, B381_kRR = 3, B381_kNM = 4, B381_kMN = 5 }; /* Ordering */

typedef struct {
  void* stream;
  b381_fr coset_gen;               /* Montgomery; one() = no coset */
  int batch_size;
  bool columns_batch;
  int ordering;
  bool are_inputs_on_device;
  bool are_outputs_on_device;
  bool is_async;
  void* ext;
} b381_ntt_config;

typedef struct {
  void* stream;
  bool is_async;
  void* ext;
} b381_ntt_init_domain_config;

typedef struct {
  void* stream;
  bool is_a_on_device;
  bool is_b_on_device;
  bool is_result_on_device;
  bool is_async;
  void* ext;
} b381_vecops_config;

b381_msm_config b381_default_msm_config(void);
b381_ntt_config b381_default_ntt_config(void);
b381_vecops_config b381_default_vecops_config(void);

/* ======================= MSM (registered: icicle_curve_api.cu:660-665) ======================= */
/* replaces msm_cuda_impl (icicle_curve_api.cu:243-407).  scalars: [batch][msm_size]; bases:
 * msm_size points (shared) or [batch][msm_size]; results: batch_size points, (x,y,1) STANDARD
 * form, identity (0,1,0). */
int b381_g1_msm(const b381_fr* scalars, const b381_g1_affine* bases, int msm_size,
                const b381_msm_config* config, b381_g1_projective* results);
/* replaces msm_g2_cuda_impl (icicle_curve_api.cu:454-618) */
int b381_g2_msm(const b381_fr* scalars, const b381_g2_affine* bases, int msm_size,
                const b381_msm_config* config, b381_g2_projective* results);
/* replace msm_precompute_bases_cuda_impl / msm_g2_precompute_bases_cuda_impl
 * (icicle_curve_api.cu:415-440, :626-650): output = precompute_factor * bases_size points,
 * block k holding 2^(k*c*ceil(W/factor)) * P_i, Montgomery affine. */
int b381_g1_msm_precompute_bases(const b381_g1_affine* input_bases, int bases_size,
                                 const b381_msm_config* config, b381_g1_affine* output_bases);
int b381_g2_msm_precompute_bases(const b381_g2_affine* input_bases, int bases_size,
                                 const b381_msm_config* config, b381_g2_affine* output_bases);

/* ======================= NTT (registered: icicle_field_api.cu:344-346) ======================= */
/* replaces ntt_init_domain_cuda_impl (src/field/ntt_kernels.cu:1607-1679).  `primitive_root` may
 * be in standard (what upstream ICICLE passes, core/ntt.rs:412-413) or Montgomery form (what the
 * reference's CUDA tests pass, tests/test_ntt_security.cu:1034-1043); its order 2^k (k<=32) is
 * discovered, not assumed. */
int b381_ntt_init_domain(const b381_fr* primitive_root, const b381_ntt_init_domain_config* config);
int b381_ntt_release_domain(void);   /* releases the CURRENT device's domain; domains are per device */
/* replaces ntt_cuda_impl (ntt_kernels.cu:968-1133) and coset_ntt_cuda_impl (:1155-1306);
 * honours ordering, coset_gen, batch_size, columns_batch and stream. size = one NTT's length. */
int b381_ntt(const b381_fr* input, int size, int dir, const b381_ntt_config* config, b381_fr* output);
/* register_ntt_get_rou_from_domain is declared (icicle_backend_api.cuh:144) but never registered
 * by the reference; provided here. Returns the Montgomery-form root of order 2^logn. */
int b381_ntt_get_rou_from_domain(uint64_t logn, b381_fr* rou);

/* ======================= vecops (registered: icicle_field_api.cu:347-352) ======================= */
/* replace vector_{add,sub,mul}_cuda_impl / scalar_{mul,add}_vec_cuda_impl
 * (icicle_field_api.cu:133-334).  For scalar_*: `scalar_a` points to ONE element. */
int b381_vector_add(const b381_fr* a, const b381_fr* b, uint64_t size, const b381_vecops_config* config, b381_fr* out);
int b381_vector_sub(const b381_fr* a, const b381_fr* b, uint64_t size, const b381_vecops_config* config, b381_fr* out);
int b381_vector_mul(const b381_fr* a, const b381_fr* b, uint64_t size, const b381_vecops_config* config, b381_fr* out);
int b381_scalar_mul_vec(const b381_fr* scalar_a, const b381_fr* b, uint64_t size, const b381_vecops_config* config, b381_fr* out);
int b381_scalar_add_vec(const b381_fr* scalar_a, const b381_fr* b, uint64_t size, const b381_vecops_config* config, b381_fr* out);
/* Upstream ICICLE v4 batch form (VecOpsConfig.batch_size / columns_batch; core/vecops.rs:345-346 sets batch_size):
 * `batch_size` scalars, `batch_size` vectors of `size` elements, stored one after the other, or interleaved when
 * columns_batch (element j of vector k at j * batch_size + k).  The ICICLE glue built with
 * -DB381_ICICLE_UPSTREAM_VECOPS routes here (csrc/icicle/field_api.cu). */
int b381_scalar_mul_vec_batch(const b381_fr* scalars, const b381_fr* b, uint64_t size, int batch_size, bool columns_batch,
                              const b381_vecops_config* config, b381_fr* out);
int b381_scalar_add_vec_batch(const b381_fr* scalars, const b381_fr* b, uint64_t size, int batch_size, bool columns_batch,
                              const b381_vecops_config* config, b381_fr* out);

/* ---- unregistered ops the reference's callers use on the same vectors ---- */
/* out[0] = sum a[i]  (vec_sum_cuda, vec_ops.cu:479-520) */
int b381_vector_sum(const b381_fr* a, uint64_t size, const b381_vecops_config* config, b381_fr* out);
/* out[i] = a[i]^-1, 0 -> 0; Montgomery's trick + one variable-time inversion per 32 elements
 * (batch_inv_cuda, vec_ops.cu:606-677) */
int b381_vector_inv(const b381_fr* a, uint64_t size, const b381_vecops_config* config, b381_fr* out);
/* out[bitrev(i)] = a[i], size a power of two; in place when out == a on the device
 * (icicle bit_reverse / bit_reverse_inplace as called by core/vecops.rs:392-535) */
int b381_bit_reverse(const b381_fr* a, uint64_t size, const b381_vecops_config* config, b381_fr* out);
/* standard <-> Montgomery form (field_to_montgomery / field_from_montgomery, field.cuh:906-928) */
int b381_montgomery_convert(const b381_fr* a, uint64_t size, int to_montgomery, const b381_vecops_config* config, b381_fr* out);

/* ======================= reference-named flat test API ======================= */
/* icicle_curve_api.cu:679-706: Montgomery points, INTEGER-form scalars unless the config flag says
 * Montgomery, DEVICE-or-host per flags, result Jacobian in Montgomery form (Z = R or 0). */
int bls12_381_g1_msm_cuda(const b381_fr* scalars, const b381_g1_affine* bases, int msm_size,
                          const b381_msm_config* config, b381_g1_projective* result);
int bls12_381_g2_msm_cuda(const b381_fr* scalars, const b381_g2_affine* bases, int msm_size,
                          const b381_msm_config* config, b381_g2_projective* result);
/* ntt_kernels.cu:1911-1942, icicle_field_api.cu:363-383 */
int bls12_381_ntt_cuda(const b381_fr* input, int size, int dir, const b381_ntt_config* config, b381_fr* output);
int bls12_381_ntt_init_domain_cuda(const b381_fr* root_of_unity, const b381_ntt_init_domain_config* config);
int bls12_381_ntt_release_domain_cuda(void);
int bls12_381_coset_ntt_cuda(const b381_fr* input, int size, int dir, const b381_fr* coset_gen,
                             const b381_ntt_config* config, b381_fr* output);
int bls12_381_field_ntt_cuda(const b381_fr* input, int size, int dir, const b381_ntt_config* config, b381_fr* output);
int bls12_381_field_ntt_init_domain_cuda(const b381_fr* root_of_unity, const b381_ntt_init_domain_config* config);
int bls12_381_field_ntt_release_domain_cuda(void);
/* vec_ops.cu:693-838 (host-or-device per config) */
int bls12_381_vector_add(const b381_fr* a, const b381_fr* b, uint64_t size, const b381_vecops_config* config, b381_fr* out);
int bls12_381_vector_sub(const b381_fr* a, const b381_fr* b, uint64_t size, const b381_vecops_config* config, b381_fr* out);
int bls12_381_vector_mul(const b381_fr* a, const b381_fr* b, uint64_t size, const b381_vecops_config* config, b381_fr* out);

/* vec_ops.cu:393-520: DEVICE vectors, output first; the scalar of scalar_*_vec_cuda is read on the HOST */
int vec_add_cuda(b381_fr* out, const b381_fr* a, const b381_fr* b, int size, const b381_vecops_config* config);
int vec_sub_cuda(b381_fr* out, const b381_fr* a, const b381_fr* b, int size, const b381_vecops_config* config);
int vec_mul_cuda(b381_fr* out, const b381_fr* a, const b381_fr* b, int size, const b381_vecops_config* config);
int scalar_mul_vec_cuda(b381_fr* out, const b381_fr* scalar, const b381_fr* vec, int size, const b381_vecops_config* config);
int scalar_add_vec_cuda(b381_fr* out, const b381_fr* scalar, const b381_fr* vec, int size, const b381_vecops_config* config);
int vec_sum_cuda(b381_fr* out, const b381_fr* a, int size, const b381_vecops_config* config);

/* ---- point formats and ingest validation (point_ops.cu:759-1000, point.cuh:339-387) ----
 * Jacobian (X, Y, Z) Montgomery, x = X/Z^2, y = Y/Z^3, infinity (0, R, 0) <-> affine Montgomery, infinity (0, 0).
 * projective_to_affine shares one inversion between 16 points (the reference inverts every Z separately).
 * Residency per config->is_a_on_device / is_result_on_device; 0 < size <= 2^26 like the reference. */
int bls12_381_g1_affine_to_projective(const b381_g1_affine* in, int size, const b381_vecops_config* config, b381_g1_projective* out);
int bls12_381_g1_projective_to_affine(const b381_g1_projective* in, int size, const b381_vecops_config* config, b381_g1_affine* out);
int bls12_381_g2_affine_to_projective(const b381_g2_affine* in, int size, const b381_vecops_config* config, b381_g2_projective* out);
int bls12_381_g2_projective_to_affine(const b381_g2_projective* in, int size, const b381_vecops_config* config, b381_g2_affine* out);
/* flags[i] = 1 if point i is infinity or satisfies y^2 = x^3 + 4 (G1) / y^2 = x^3 + 4(1+u) (G2), else 0 */
int b381_g1_is_on_curve(const b381_g1_affine* in, int size, const b381_vecops_config* config, uint8_t* flags);
int b381_g2_is_on_curve(const b381_g2_affine* in, int size, const b381_vecops_config* config, uint8_t* flags);
/* flags[i] = 1 if point i (assumed on the curve) lies in the order-r subgroup: phi(P) = [z^2 - 1] P on G1,
 * psi(P) = [z] P on G2 (csrc/glv.cuh); infinity counts as a member.  The reference leaves both checks as TODO
 * (include/point.cuh:419-448). */
int b381_g1_is_in_subgroup(const b381_g1_affine* in, int size, const b381_vecops_config* config, uint8_t* flags);
int b381_g2_is_in_subgroup(const b381_g2_affine* in, int size, const b381_vecops_config* config, uint8_t* flags);
/* out[i] = scalars[i] * bases[i]  (bls12-381/src/curve/point_ops.cu:1019-1268; same names, argument order and codes):
 * bases Montgomery affine (is_a_on_device), scalars CANONICAL integers (is_b_on_device), out Jacobian Montgomery,
 * always the normalised representative (x, y, 1) / (0, R, 0).  _glv splits k = k1 + k2 * lambda with the G1
 * endomorphism (128 doublings instead of 252); the plain variant is its comparator. */
int bls12_381_g1_scalar_mul_glv(const b381_g1_affine* bases, const b381_fr* scalars, int size, const b381_vecops_config* config, b381_g1_projective* out);
int bls12_381_g1_scalar_mul(const b381_g1_affine* bases, const b381_fr* scalars, int size, const b381_vecops_config* config, b381_g1_projective* out);

/* ======================= device plumbing (CudaDeviceAPI, src/device/cuda_device_api.cu:38-149) ==== */
int b381_device_count(int* count);
int b381_set_device(int device_id);
int b381_malloc(void** ptr, size_t size);
int b381_malloc_async(void** ptr, size_t size, void* stream);
int b381_free(void* ptr);
int b381_free_async(void* ptr, void* stream);
int b381_memset(void* ptr, int value, size_t size);
int b381_copy_to_device(void* dst, const void* src, size_t size);
int b381_copy_to_host(void* dst, const void* src, size_t size);
int b381_copy_to_device_async(void* dst, const void* src, size_t size, void* stream);
int b381_copy_to_host_async(void* dst, const void* src, size_t size, void* stream);
int b381_copy_device_to_device(void* dst, const void* src, size_t size, void* stream);
int b381_host_alloc_pinned(void** ptr, size_t size);
int b381_host_free_pinned(void* ptr);
int b381_stream_create(void** stream);
int b381_stream_destroy(void* stream);
int b381_stream_synchronize(void* stream);
int b381_device_synchronize(void);

/* ======================= multi-GPU building blocks (north_star: point-range sharding) ========== */
/* One rank's share of an MSM: same as b381_g1_msm but leaves the partial sum as an XYZZ point in
 * Montgomery form on the device (4 x Fq = 192 B; 4 x Fq2 = 384 B for G2), no inversion. */
int b381_g1_msm_partial(const b381_fr* scalars, const b381_g1_affine* bases, int msm_size,
                        const b381_msm_config* config, void* partial_xyzz_device);
int b381_g2_msm_partial(const b381_fr* scalars, const b381_g2_affine* bases, int msm_size,
                        const b381_msm_config* config, void* partial_xyzz_device);
/* Distributed four-step NTT, step 1 (north_star: "large NTTs use the four-step decomposition with an
 * NCCL all-to-all").  A 2^log_n transform is spread over 2^log_gpus GPUs by COLUMN BLOCKS: with
 * lo = log_n - upper_stages, GPU `rank` holds x[(i_hi << lo) | l] for every i_hi and l in
 * [rank*L, (rank+1)*L), L = 2^lo / #GPUs, stored locally as [i_hi][l - rank*L].  This call runs the
 * top `upper_stages` DIF stages in place (twiddles follow the global index).  The caller then
 * exchanges row blocks (all-to-all) and finishes with b381_ntt(batch = 2^upper_stages / #GPUs,
 * size = 2^lo, ordering kNR) -- see midnight_bls12_381_cuda_b200/dist.py. */
int b381_ntt_dist_columns(b381_fr* data_device, int log_n, int log_gpus, int rank, int upper_stages, int dir, void* stream);
/* Same, with the exchange FUSED into the last column pass: its stores go straight into the row buffers of the
 * owning GPUs (`peer_rows[r]` = rank r's [2^upper_stages / #GPUs][2^lo] buffer mapped into this process, e.g. with
 * b381_ipc_alloc / b381_ipc_open), already transposed, so no all-to-all and no transpose pass follow: the caller
 * only needs a barrier across ranks before the row transforms (and before the next call reuses the buffers).
 * At most 8 GPUs.  `data_device` is scratch afterwards. */
int b381_ntt_dist_columns_p2p(b381_fr* data_device, int log_n, int log_gpus, int rank, int upper_stages, int dir,
                              void* const* peer_rows, void* stream);
/* Peer-visible device memory for the call above (one process per GPU): cudaMalloc + its 64-byte CUDA IPC handle;
 * b381_ipc_open maps a peer's handle into this process (peer access is enabled lazily), b381_ipc_close unmaps it;
 * the owner releases the allocation with b381_free. */
int b381_ipc_alloc(size_t bytes, void** ptr, unsigned char handle[64]);
int b381_ipc_open(const unsigned char handle[64], void** ptr);
int b381_ipc_close(void* ptr);
/* Adds `count` XYZZ partials (device) and writes ONE ICICLE standard-form projective result.  result_on_device = false:
 * synchronous, the result has landed in host memory on return; true: stream-ordered on `stream`, no host synchronisation. */
int b381_g1_msm_combine(const void* partials_xyzz_device, int count, void* stream, bool result_on_device,
                        b381_g1_projective* result);
int b381_g2_msm_combine(const void* partials_xyzz_device, int count, void* stream, bool result_on_device,
                        b381_g2_projective* result);

/* ======================= base-set generation ======================= */
/* out_device[i] = P0 + i*D as Montgomery affine points (P0, D: HOST pointers to one point each; out:
 * DEVICE).  Lays down large distinct base sets with known discrete logs; conversion to affine is
 * batched (one inversion per 16 points).  Reference counterpart: the per-point conversions of
 * bls12-381/src/curve/point_ops.cu:61-101. */
int b381_g1_point_series(const b381_g1_affine* p0, const b381_g1_affine* d, uint64_t n, b381_g1_affine* out_device, void* stream);
int b381_g2_point_series(const b381_g2_affine* p0, const b381_g2_affine* d, uint64_t n, b381_g2_affine* out_device, void* stream);

/* ======================= measurement helpers (bench.py / profiles) ======================= */
/* Dependent-free IMAD.WIDE.U32 issue-rate probe: returns MAD/s through *mads_per_s. */
int b381_bench_imad_peak(int iters, double* mads_per_s, float* ms);
/* Back-to-back Montgomery multiplications per thread (throughput of fq/fr mul in isolation). */
int b381_bench_field_mul(int field /*0=fq,1=fr*/, int iters, double* muls_per_s, float* ms);
/* last kernel-level timing breakdown of the most recent MSM on this thread (B381_MSM_TIMING=1), ms per phase:
 * [digits, sort, offsets, affine pre-reduction, tasks+accumulate, finalize, reduce, combine]; returns count written. */
int b381_msm_last_timings(float* out, int cap);
/* shape of the most recent MSM on this thread: [window bits c, windows W, affine pre-reduction levels,
 * own kernel launches]; returns count written. */
int b381_msm_last_info(int* out, int cap);
/* duration of the two dominant kernels of the most recent G1 MSM on this thread (level-0 forward and backward pass
 * of the affine pre-reduction), CUDA events on the launching stream, B381_MSM_TIMING=1; returns 1 if available. */
int b381_msm_last_level0_ms(float* fwd_ms, float* bwd_ms);
/* shape of the most recent NTT on this thread: [kernel launches (passes), log2 size, batch, 0]; returns count written. */
int b381_ntt_last_info(int* out, int cap);
const char* b381_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B381_H */
