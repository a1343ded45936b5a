// Exercises include/b381.hpp (the C++ twin of the reference's Rust core/ API).
//   cpp_api_test link   -> no GPU needed: layouts, default configs, error path without a device
//   cpp_api_test gpu    -> 5*G (core/msm.rs:1667-1678), sum i*G = 2080 G (:1681-1694), NTT(delta) = ones
//                          (core/ntt.rs:2059-2073), round trip, vector_mul identity; prints "cpp api ok"
#include <cstdio>
#include <cstring>
#include <string>

#include "b381.hpp"
using namespace b381;

static const uint64_t G1X[6] = {0x5cb38790fd530c16ull, 0x7817fc679976fff5ull, 0x154f95c7143ba1c1ull, 0xf0ae6acdf3d0e747ull, 0xedce6ecc21dbf440ull, 0x120177419e0bfb75ull};
static const uint64_t G1Y[6] = {0xbaac93d50ce72271ull, 0x8c22631a7918fd8eull, 0xdd595f13570725ceull, 0x51ac582950405194ull, 0x0e1c8c3fad0059c0ull, 0x0bbc3efc5008a26aull};
static const uint64_t FR_ONE[4] = {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full};
// 7^((r-1)/2^32)^(2^(32-10)) is not needed: pass the 2^32-th root (Montgomery), the backend discovers the order... too big a
// table; use omega_{2^10} = ROOT^(2^22) computed by repeated vector_mul on the device instead.
static const uint64_t FR_ROOT32[4] = {0xb9b58d8c5f0e466aull, 0x5b1b4c801819d7ecull, 0x0af53ae352a31e64ull, 0x5bf3adda19e9b27bull};

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
  std::string mode = argc > 1 ? argv[1] : "link";
  b381_msm_config mc = b381_default_msm_config();
  REQUIRE(mc.batch_size == 1 && mc.precompute_factor == 1 && mc.are_points_shared_in_batch);
  b381_ntt_config nc = b381_default_ntt_config();
  REQUIRE(std::memcmp(&nc.coset_gen, FR_ONE, 32) == 0 && nc.ordering == B381_kNN);
  if (mode == "link") {
    // without a device every compute call must FAIL LOUDLY (error code -> exception), never fall back
    if (!is_gpu_available()) {
      bool threw = false;
      try {
        GpuMsmContext ctx;
        Scalar s{}; G1Affine g{};
        std::memcpy(&g.x, G1X, 48); std::memcpy(&g.y, G1Y, 48);
        ctx.msm({s}, {g});
      } catch (const Error&) { threw = true; }
      REQUIRE(threw);
    }
    std::puts("cpp api link ok");
    return 0;
  }
  G1Affine g{};
  std::memcpy(&g.x, G1X, 48); std::memcpy(&g.y, G1Y, 48);
  Scalar one; std::memcpy(&one, FR_ONE, 32);
  GpuMsmContext ctx;
  // 5*G vs G+G+G+G+G (five Montgomery ones on five copies of G)
  Scalar five{};
  { std::vector<Scalar> ones(5, one); five = vecops::scalar_add(one, vecops::scalar_add(one, vecops::scalar_add(one, vecops::scalar_add(one, {one}))))[0]; }
  G1Projective a = ctx.msm({five}, {g});
  G1Projective b = ctx.msm(std::vector<Scalar>(5, one), std::vector<G1Affine>(5, g));
  REQUIRE(std::memcmp(&a, &b, sizeof(a)) == 0 && a.z.l[0] == 1);
  // sum_{i=1..64} i*G == 2080*G, via device-resident bases, sync and async
  std::vector<Scalar> sc; Scalar acc = one;
  for (int i = 0; i < 64; i++) { sc.push_back(acc); acc = vecops::scalar_add(one, {acc})[0]; }
  Scalar k2080{}; { Scalar t{}; bool first = true; for (auto& s : sc) { t = first ? s : vecops::vector_add({t}, {s})[0]; first = false; } k2080 = t; }
  PrecomputedBases dev = ctx.upload_g1_bases(std::vector<G1Affine>(64, g));
  G1Projective s1 = ctx.msm_with_device_bases(sc, dev);
  G1Projective s2 = ctx.msm({k2080}, {g});
  REQUIRE(std::memcmp(&s1, &s2, sizeof(s1)) == 0);
  G1Projective s3 = ctx.msm_with_device_bases_async(sc, dev).wait();
  REQUIRE(std::memcmp(&s1, &s3, sizeof(s1)) == 0);
  PrecomputedBases pre = ctx.precompute_bases(dev, 4);
  G1Projective s4 = ctx.msm_with_device_bases(sc, pre);
  REQUIRE(std::memcmp(&s1, &s4, sizeof(s1)) == 0);
  // NTT: omega_{2^10} = ROOT32^(2^22) by 22 squarings on the device
  Scalar w; std::memcpy(&w, FR_ROOT32, 32);
  for (int i = 0; i < 22; i++) w = vecops::vector_mul({w}, {w})[0];
  GpuNttContext ntt(10, w);
  std::vector<Scalar> delta(1024, Scalar{});
  delta[0] = one;
  std::vector<Scalar> y = ntt.forward_ntt(delta);
  for (auto& e : y) REQUIRE(std::memcmp(&e, &one, 32) == 0);
  std::vector<Scalar> back = ntt.inverse_ntt(y);
  REQUIRE(std::memcmp(back.data(), delta.data(), 1024 * 32) == 0);
  std::puts("cpp api ok");
  return 0;
}
