// C ABI of the G1 MSM (include/b381.h) + the group-independent entry points.  Templates in msm_impl.cuh.
#include "msm_impl.cuh"

using namespace b381;

static_assert(sizeof(b381_g1_affine) == sizeof(g1_affine) && sizeof(b381_g1_projective) == sizeof(g1_jac), "layout");
static_assert(sizeof(b381_fr) == sizeof(fr_t), "layout");

// The reference's flat entry points call msm::msm_cuda directly, which never reads the Montgomery
// flags: points are taken as Montgomery, scalars as integers (icicle_curve_api.cu:679-706,
// msm_kernels.cu:603-903).  Same contract here, whatever the flags say.
static b381_msm_config flat_cfg(const b381_msm_config* cfg) {
  b381_msm_config c = *cfg;
  c.are_points_montgomery_form = true;
  c.are_scalars_montgomery_form = false;
  c.batch_size = 1;
  c.precompute_factor = 1;
  return c;
}

extern "C" {

int b381_g1_msm(const b381_fr* s, const b381_g1_affine* p, int n, const b381_msm_config* cfg, b381_g1_projective* r) {
  return msm_entry<fq_t>((const fr_t*)s, (const g1_affine*)p, n, cfg, r, ResultKind::IcicleStd);
}
int bls12_381_g1_msm_cuda(const b381_fr* s, const b381_g1_affine* p, int n, const b381_msm_config* cfg,
                          b381_g1_projective* r) {
  if (!cfg) return B381_INVALID_POINTER;
  b381_msm_config c = flat_cfg(cfg);
  return msm_entry<fq_t>((const fr_t*)s, (const g1_affine*)p, n, &c, r, ResultKind::JacobianMont);
}
int b381_g1_msm_partial(const b381_fr* s, const b381_g1_affine* p, int n, const b381_msm_config* cfg, void* out) {
  return msm_entry<fq_t>((const fr_t*)s, (const g1_affine*)p, n, cfg, out, ResultKind::PartialXyzz);
}
int b381_g1_msm_combine(const void* parts, int count, void* stream, bool on_device, b381_g1_projective* r) {
  return combine_entry<fq_t>(parts, count, stream, on_device, r);
}
int b381_g1_msm_precompute_bases(const b381_g1_affine* in, int n, const b381_msm_config* cfg, b381_g1_affine* out) {
  return precompute_entry<fq_t>((const g1_affine*)in, n, cfg, (g1_affine*)out);
}
int b381_msm_last_timings(float* out, int cap) {
  int k = g_last_timings_n < cap ? g_last_timings_n : cap;
  for (int i = 0; i < k; i++) out[i] = g_last_timings[i];
  return k;
}
int b381_msm_last_info(int* out, int cap) {
  int k = cap < 4 ? cap : 4;
  for (int i = 0; i < k; i++) out[i] = g_last_info[i];
  return k;
}
b381_msm_config b381_default_msm_config(void) {
  b381_msm_config c;
  memset(&c, 0, sizeof(c));
  c.precompute_factor = 1;
  c.batch_size = 1;
  c.are_points_shared_in_batch = true;
  return c;
}

}  // extern "C"
