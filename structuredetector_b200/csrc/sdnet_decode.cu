// sdnet_decode.cu -- B200 (sm_100a) kernels + C ABI for the SDNet decoding path.
//
// Replaces the tensor half of the reference decoder
//   src/sdnet/data/decoders.py:44-100  (Decoder.__call__)
//   src/sdnet/utils/utils.py:355-361   clamped_sigmoid
//   src/sdnet/utils/utils.py:441-443   nms (5x5 max-pool equality)
//   src/sdnet/utils/utils.py:447-467   topk (per-channel, then cross-channel)
//   src/sdnet/utils/utils.py:347-351   transpose_and_gather
//   src/sdnet/utils/utils.py:422-437   hypot
// with three launches:
//   1. peaks kernel   -- the only pass over the heat maps (HBM-bound).  One warp walks a
//      128-column panel of one plane top to bottom with a register ring of 8 rows, finds
//      the pixels that survive the reference's NMS and appends (score, index) records to a
//      small per-plane candidate list, pruning everything that provably cannot reach the
//      plane's top-K.
//   2. exact-select kernel -- only for planes whose candidate list overflowed (huge exact
//      plateaus): bounded-memory radix select straight from the heat map.
//   3. tail kernel    -- one CTA per image: radix-select + sort of the candidates under the
//      total order (score desc, class asc, index asc), zero-fill, offset/embedding gather,
//      coordinate assembly, masking, nearest-anchor grouping.
//
// Numerics contract:
//   * score  = min(max(1/(1+expf(-x)), 1e-6f), (float)(1-1e-6))   -- ATen's CUDA formula;
//   * a pixel survives NMS iff score == max score over its window.  With S monotone
//     non-decreasing in x (verified exhaustively on the device, tests/test_gpu_parity.py)
//     this is  S(x) == S(window-max of x), so the stencil runs on raw logits and the exact
//     sigmoid is evaluated only for the few pixels within a hair of their window maximum;
//   * ties: (score desc, class asc, flat index asc) -- what torch.topk does on CUDA for k > 32;
//   * grouping arithmetic uses explicitly rounded mul/add/sqrt (no FMA contraction) and the
//     first minimum wins.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdlib.h>

#include "sdnet_decode.h"

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int kThreads = 256;          // peaks kernel CTA
constexpr int kWarps = kThreads / 32;
constexpr int kPanelW = 128;           // columns per warp (32 lanes x 4)
constexpr int kBins = 128;             // per-warp logit histogram used for pruning
constexpr float kBinLo = -16.0f;
constexpr float kBinScale = 4.0f;      // bins of 0.25 logit
constexpr int kFineBins = 512;         // shared (CTA-wide / plane-wide) histograms: bins of 1/16 logit
constexpr float kFineScale = 16.0f;    // (256 x 0.125: blobs 3 % slower; 1024 x 1/32: flushes 25 % slower)
constexpr int kFinePerLane = kFineBins / 32;
constexpr float kSatX = 14.0f;         // |x| >= 14 is inside the clamp on both sides: S(x) == S(+-14)
constexpr float kPreScale = 8.0f;      // pre-activated maps: histogram runs on 8*value
constexpr float kClampLo = 1e-6f;
constexpr float kClampHi = (float)(1.0 - 1e-6);
constexpr float kFar = 1e6f;

constexpr int kSortN = 2048;           // tail sort buffer (>= SDNET_MAX_TOPK + boundary slack)

struct View4 {
  const void* data;      // element type given by the launch's dtype
  long long sb, sc, sh;  // strides in elements
};

struct PeaksParams {
  View4 anchor, part;
  int B, M, N, H, W, K, P;
  int strips, rows_per_strip, panels;
  int units;
  int cap;                 // records per plane list
  int pre_activated;
  u64* lists;              // [planes][cap]
  int* counts;             // [planes] records emitted (may exceed cap)
  u32* sched;              // [0] dynamic unit counter
  u32* ghist;              // [planes][kFineBins] plane-wide logit histogram of recorded candidates
  int* gfloor;             // [planes] highest fine bin b with >= K recorded candidates in bins >= b (0 = none)
  int l2_prefetch_groups;  // warp-specialised kernel: L2 prefetch distance in 4-row groups (0 = off)
  // tile kernel, two-tier schedule: units [0, tier1_units) are whole-height panels of planes
  // [0, tier1_planes); the remaining planes are cut into `strips` strips so that the last wave of
  // warps is filled with short units instead of idling behind a few long ones
  int tier1_units, tier1_planes;
  int odd_x;               // tile kernel, row-pair maps: tensor-map x coordinate of an odd row's column 0
};

// Numerics of the score function per input dtype DT (SDNET_DTYPE_*).
//
// fp32: S(x) = clamp(1/(1+expf(-x))), bit-identical to ATen's CUDA kernels (UnarySpecialOpsKernel.cu
// sigmoid: one / (one + std::exp(-a)); TensorCompare.cu clamp: min(max(v, lo), hi)).
// fp16 / bf16 (what the reference's `--amp` validation feeds the decoder): ATen evaluates both ops in
// fp32 and rounds each result to the tensor dtype, so S_T(x) = T(clamp(float(T(sigmoid(float(x)))))),
// returned here as the exactly representable float.  Every S_T is monotone non-decreasing in x.
//
// The margins say when two different logits x < h might share a score: only if x >= h - kNear with h in
// [kLo, kHi], or x >= h - kNear2 with h in (kHi, kHi2], or h > kHi2 and x > kHi2 - 1, or h < kLo.
// Outside that, S(x) < S(h) strictly; verified exhaustively on the device for every dtype
// (tests/test_gpu_parity.py, tests/test_gpu_halfprec.py).
template <int DT>
struct Num;

template <>
struct Num<SDNET_DTYPE_F32> {
  typedef float In;
  static constexpr float kNear = 2e-3f, kHi = 8.0f, kLo = -13.0f;
  static constexpr float kNear2 = 2e-3f, kHi2 = 8.0f;  // no second zone
  static __device__ __forceinline__ float act(float x) {
    const float s = 1.0f / (1.0f + expf(-x));
    return fminf(fmaxf(s, kClampLo), kClampHi);
  }
  static __device__ __forceinline__ float to_float(float v) { return v; }
};

template <>
struct Num<SDNET_DTYPE_F16> {
  typedef __half In;
  static constexpr float kNear = 0.02f, kHi = 3.0f, kLo = -11.0f;
  static constexpr float kNear2 = 0.15f, kHi2 = 5.0f;  // 10-bit mantissa: ties reach 0.073 logit at h = 5
  static __device__ __forceinline__ float act(float x) {
    const float s = __half2float(__float2half_rn(1.0f / (1.0f + expf(-x))));
    return __half2float(__float2half_rn(fminf(fmaxf(s, kClampLo), kClampHi)));
  }
  static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
};

template <>
struct Num<SDNET_DTYPE_BF16> {
  typedef __nv_bfloat16 In;
  static constexpr float kNear = 0.1f, kHi = 2.0f, kLo = -13.0f;
  static constexpr float kNear2 = 0.6f, kHi2 = 4.0f;   // 7-bit mantissa: ties reach 0.22 logit at h = 4
  static __device__ __forceinline__ float act(float x) {
    const float s = __bfloat162float(__float2bfloat16_rn(1.0f / (1.0f + expf(-x))));
    return __bfloat162float(__float2bfloat16_rn(fminf(fmaxf(s, kClampLo), kClampHi)));
  }
  static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
};

// element `idx` of a tensor whose dtype is DT, as float (exact)
template <int DT>
__device__ __forceinline__ float ld_in(const void* base, long long idx) {
  return Num<DT>::to_float(__ldg(static_cast<const typename Num<DT>::In*>(base) + idx));
}

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// Programmatic dependent launch: the next kernel of the decode may be scheduled while this one is
// still running (its CTAs take whatever SM resources free up and park at pdl_wait), which hides the
// launch latency between the three kernels.  pdl_wait returns once the previous kernel has fully
// completed and its memory is visible.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float comp(const float4& v, int j) {
  return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
}

// ---------------------------------------------------------------------------------------------
// peaks kernel
//
// Work unit = (plane, row strip, 128-column panel), one warp per unit, units handed out by an
// atomic counter.  Each warp streams its panel top to bottom through a private shared-memory
// ring of kStages rows filled with cp.async (16 B per lane, straight from L2, no registers):
//     row buffer (kPitch floats):  [pad pad hL hL | 128 panel columns | hR hR pad pad]
// Per row the common case is: issue the copy of the row kStages-1 ahead, wait for the row two
// below the centre, read the centre row (one LDS.128), and vote "does any pixel beat the
// pruning floor?".  Only then is the 5x5 window maximum formed (vertical max from the ring,
// neighbours' columns by shuffle, panel-edge columns from the halo slots) and the exact
// sigmoid evaluated for the pixels within a hair of their window maximum.
// ---------------------------------------------------------------------------------------------
constexpr int kStages = 8;             // ring depth (power of two); kStages - 1 - 2R rows stay in flight
constexpr int kPitch = 136;            // floats per ring row
constexpr int kPitchB = kPitch * 4;
constexpr int kBuf = 64;               // per-warp candidate buffer (records), flushed at >= 32
constexpr int kPeaksSmemPerWarp = kStages * kPitchB + kBins * 8 + kBuf * 8 + kStages * 8;
constexpr int kPeaksSmem = kWarps * kPeaksSmemPerWarp;

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
// predicated cp.async: global -> shared, no register staging
__device__ __forceinline__ void cp_async4_if(u32 dst, const void* src, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 4;\n\t}"
               ::"r"(dst), "l"(src), "r"((int)pred));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds128(u32 addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64(u32 addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(u32 addr, float a) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ void sts64(u32 addr, float a) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %1};" ::"r"(addr), "f"(a) : "memory");
}

// ---- mbarrier + 1-D bulk copy (TMA unit; SASS: UBLKCP, SYNCS) ---------------------------------
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(u32 bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#ifndef SDNET_X_WAIT
#define SDNET_X_WAIT 0
#endif
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
#if SDNET_X_WAIT == 0
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@!p bra WAIT_LOOP;\n\t}"
      ::"r"(bar), "r"(parity), "r"(1000u) : "memory");  // suspend-time hint (ns): sleep instead of spinning
#elif SDNET_X_WAIT == 1
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_LOOP;\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "WAIT_LOOP:\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_LOOP;\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, u32 bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Feeds one warp's ring in the fallback kernel: every lane copies its own four columns with
// 4-byte cp.async (any alignment).  Ring sequence number q of a unit <-> image row row0 + q.
template <bool kBulk>
struct RowFeed;

template <>
struct RowFeed<false> {
  u32 ring_s, base;
  const char* gown;
  const char* ghalo;
  long long pitch;
  u32 s_own, s_halo, own_ok, halo_ok;
  int row0, H, q_last;

  __device__ __forceinline__ void init(u32 ring, u32, int lane) {
    ring_s = ring; base = 0;
    s_own = ring + (4 + 4 * lane) * 4;
    s_halo = ring + (lane == 31 ? 4 + kPanelW : 2) * 4;
  }
  __device__ __forceinline__ void begin_unit(const void* plane_v, long long sh, int r0, int H_, int W, int panel_col0,
                                             int q_last_, int lane) {
    const float* plane = static_cast<const float*>(plane_v);
    row0 = r0; H = H_; q_last = q_last_;
    const int col0 = panel_col0 + 4 * lane;
    const int halo_col = lane == 31 ? panel_col0 + kPanelW : panel_col0 - 2;
    pitch = sh * 4;
    gown = reinterpret_cast<const char*>(plane + (long long)r0 * sh + col0);
    ghalo = reinterpret_cast<const char*>(plane + (long long)r0 * sh + halo_col);
    own_ok = 0;
    for (int jj = 0; jj < 4; ++jj) own_ok |= (col0 + jj < W ? 1u : 0u) << jj;
    halo_ok = 0;
    if (lane == 0 || lane == 31)
      for (int jj = 0; jj < 2; ++jj) halo_ok |= ((halo_col + jj >= 0 && halo_col + jj < W) ? 1u : 0u) << jj;
    for (int i = lane; i < kStages * kPitch / 4; i += 32) sts128(ring_s + 16 * i, -CUDART_INF_F);
    __syncwarp();
  }
  __device__ __forceinline__ u32 slot_addr(int q) const { return ring_s + (q & (kStages - 1)) * kPitchB; }
  __device__ __forceinline__ void issue(int q, int lane) {
    if (q <= q_last) {
      const u32 so = s_own + (q & (kStages - 1)) * kPitchB, sh = s_halo + (q & (kStages - 1)) * kPitchB;
      if ((unsigned)(row0 + q) < (unsigned)H) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) cp_async4_if(so + 4 * jj, gown + 4 * jj, (own_ok >> jj) & 1u);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) cp_async4_if(sh + 4 * jj, ghalo + 4 * jj, (halo_ok >> jj) & 1u);
      } else {
        sts128(so, -CUDART_INF_F);
        if (lane == 0 || lane == 31) sts64(sh, -CUDART_INF_F);
      }
    }
    gown += pitch;
    ghalo += pitch;
    cp_async_commit();
  }
  // cp.async groups complete in order: allowing kStages-1-2R groups in flight means the row
  // two below the centre has landed
  template <int R>
  __device__ __forceinline__ void wait_step() const { cp_async_wait<kStages - 1 - 2 * R>(); }
  __device__ __forceinline__ void end_unit() { cp_async_wait<0>(); }
};

// Feed for fp16 / bf16 maps: the ring stays fp32 (so everything downstream is shared with the
// fp32 path); each lane loads its own four columns (and lanes 0 / 31 the two halo columns) with
// plain 2-byte loads -- any alignment, any W -- converts, and stores to the ring three steps
// later, so kDepth rows per warp are in flight in registers.  Same call pattern as RowFeed<false>:
// issue(q) makes row q - kDepth resident, which is exactly the row step q - 7 needs.
template <int DT>
struct RowFeedCvt {
  typedef typename Num<DT>::In In;
  static constexpr int kDepth = 3;
  u32 ring_s, s_own, s_halo, own_ok, halo_ok;
  bool vec_ok;  // every row of this lane's four columns is 8-byte aligned
  const In* gown;
  const In* ghalo;
  long long pitch;
  int row0, H, q_last;
  float4 own[kDepth];
  float2 halo[kDepth];

  __device__ __forceinline__ void init(u32 ring, u32, int lane) {
    ring_s = ring;
    s_own = ring + (4 + 4 * lane) * 4;
    s_halo = ring + (lane == 31 ? 4 + kPanelW : 2) * 4;
  }
  __device__ __forceinline__ void begin_unit(const void* plane_v, long long sh, int r0, int H_, int W, int panel_col0,
                                             int q_last_, int lane) {
    const In* plane = static_cast<const In*>(plane_v);
    row0 = r0; H = H_; q_last = q_last_;
    const int col0 = panel_col0 + 4 * lane;
    const int halo_col = lane == 31 ? panel_col0 + kPanelW : panel_col0 - 2;
    pitch = sh;
    gown = plane + (long long)r0 * sh + col0;
    ghalo = plane + (long long)r0 * sh + halo_col;
    vec_ok = (reinterpret_cast<uintptr_t>(gown) % 8 == 0) && (sh % 4 == 0);
    own_ok = 0;
    for (int jj = 0; jj < 4; ++jj) own_ok |= (col0 + jj < W ? 1u : 0u) << jj;
    halo_ok = 0;
    if (lane == 0 || lane == 31)
      for (int jj = 0; jj < 2; ++jj) halo_ok |= ((halo_col + jj >= 0 && halo_col + jj < W) ? 1u : 0u) << jj;
    for (int i = lane; i < kStages * kPitch / 4; i += 32) sts128(ring_s + 16 * i, -CUDART_INF_F);
    __syncwarp();
  }
  __device__ __forceinline__ u32 slot_addr(int q) const { return ring_s + (q & (kStages - 1)) * kPitchB; }
  __device__ __forceinline__ void issue(int q, int lane) {
    const float ninf = -CUDART_INF_F;
    if (q >= kDepth && q - kDepth <= q_last) {  // retire the oldest register stage into the ring
      const u32 slot = (u32)(q - kDepth) & (kStages - 1);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(s_own + slot * kPitchB), "f"(own[0].x), "f"(own[0].y),
                   "f"(own[0].z), "f"(own[0].w) : "memory");
      if (lane == 0 || lane == 31)
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(s_halo + slot * kPitchB), "f"(halo[0].x), "f"(halo[0].y) : "memory");
    }
#pragma unroll
    for (int d = 0; d + 1 < kDepth; ++d) { own[d] = own[d + 1]; halo[d] = halo[d + 1]; }
    float4 v = make_float4(ninf, ninf, ninf, ninf);
    float2 hv = make_float2(ninf, ninf);
    if (q <= q_last && (unsigned)(row0 + q) < (unsigned)H) {
      if (own_ok == 15u && vec_ok) {  // four columns in one 8-byte load
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(gown));
        In e[4];
        memcpy(e, &raw, 8);
        v = make_float4(Num<DT>::to_float(e[0]), Num<DT>::to_float(e[1]), Num<DT>::to_float(e[2]), Num<DT>::to_float(e[3]));
      } else {
        if (own_ok & 1u) v.x = Num<DT>::to_float(__ldg(gown + 0));
        if (own_ok & 2u) v.y = Num<DT>::to_float(__ldg(gown + 1));
        if (own_ok & 4u) v.z = Num<DT>::to_float(__ldg(gown + 2));
        if (own_ok & 8u) v.w = Num<DT>::to_float(__ldg(gown + 3));
      }
      if (halo_ok & 1u) hv.x = Num<DT>::to_float(__ldg(ghalo + 0));
      if (halo_ok & 2u) hv.y = Num<DT>::to_float(__ldg(ghalo + 1));
    }
    own[kDepth - 1] = v;
    halo[kDepth - 1] = hv;
    gown += pitch;
    ghalo += pitch;
  }
  template <int R>
  __device__ __forceinline__ void wait_step() const {}
  __device__ __forceinline__ void end_unit() {}
};

template <int DT>
struct FeedFor { typedef RowFeedCvt<DT> type; };
template <>
struct FeedFor<SDNET_DTYPE_F32> { typedef RowFeed<false> type; };

__device__ __forceinline__ int logit_bin(float x) {
  int bin = __float2int_rd((x - kBinLo) * kBinScale);
  bin = max(0, min(kBins - 1, bin));
  // rounding guard: never count an element in a bin whose lower edge is above it
  if (bin > 0 && x < kBinLo + (float)bin * (1.0f / kBinScale)) --bin;
  return bin;
}

// order-preserving float <-> int (so atomicMin works on floats of either sign)
__device__ __forceinline__ int ord_of(float x) {
  const int b = __float_as_int(x);
  return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord_to_float(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7fffffff)); }

// Highest bin b with (count in bins >= b) >= K given each lane's four bin counts; -1 if none.
__device__ __forceinline__ int floor_bin_of(const uint4 c, int lane, int K) {
  const u32 s = c.x + c.y + c.z + c.w;
  u32 suf = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 t = __shfl_down_sync(0xffffffffu, suf, d);
    if (lane + d < 32) suf += t;
  }
  const u32 mask = __ballot_sync(0xffffffffu, suf >= (u32)K);
  if (mask == 0) return -1;
  const int L = 31 - __clz(mask);
  int b = 0;
  if (lane == L) {
    u32 above = suf - s;
    if (above + c.w >= (u32)K) b = 4 * L + 3;
    else if (above + c.w + c.z >= (u32)K) b = 4 * L + 2;
    else if (above + c.w + c.z + c.y >= (u32)K) b = 4 * L + 1;
    else b = 4 * L;
  }
  return __shfl_sync(0xffffffffu, b, L);
}

// Warp-local pruning floor.  Let b be the highest bin such that this warp-unit has already
// recorded >= K candidates in bins >= b.  Every one of those has a (saturation-clamped) logit
// >= minx[b] and a lower flat index than anything the unit will see later, so it beats any
// later pixel whose clamped logit is <= minx[b] under (score desc, index asc) -- equal scores
// included.  Returns minx[b], or -inf when fewer than K candidates were recorded.
__device__ __forceinline__ float local_floor(const u32* hist, const int* minx, int lane, int K) {
  const int b = floor_bin_of(*reinterpret_cast<const uint4*>(hist + 4 * lane), lane, K);
  if (b < 0) return -CUDART_INF_F;
  return ord_to_float(minx[b]);
}

__device__ __forceinline__ int fine_bin(float x) {
  int bin = __float2int_rd((x - kBinLo) * kFineScale);
  bin = max(0, min(kFineBins - 1, bin));
  if (bin > 0 && x < kBinLo + (float)bin * (1.0f / kFineScale)) --bin;  // rounding guard, as in logit_bin
  return bin;
}

// Highest fine bin b with (count in bins >= b) >= K; lane l owns bins 32l..32l+31.  -1 if none.
// kShared: the histogram lives in shared memory (volatile loads) instead of global (L2 loads).
template <bool kShared>
__device__ __forceinline__ uint4 load_bins(const u32* ptr) {
  uint4 v;
  if (kShared) {
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(ptr)));
  } else {
    v = __ldcg(reinterpret_cast<const uint4*>(ptr));
  }
  return v;
}

template <bool kShared>
__device__ __forceinline__ int floor_bin_fine(const u32* hist, int lane, int K) {
  const u32* mine = hist + kFinePerLane * lane;
  uint4 v[kFinePerLane / 4];
#pragma unroll
  for (int q = 0; q < kFinePerLane / 4; ++q) v[q] = load_bins<kShared>(mine + 4 * q);
  u32 s = 0;
#pragma unroll
  for (int q = 0; q < kFinePerLane / 4; ++q) s += v[q].x + v[q].y + v[q].z + v[q].w;
  u32 suf = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 t = __shfl_down_sync(0xffffffffu, suf, d);
    if (lane + d < 32) suf += t;
  }
  const u32 mask = __ballot_sync(0xffffffffu, suf >= (u32)K);
  if (mask == 0) return -1;
  const int L = 31 - __clz(mask);
  int b = kFinePerLane * L;
  if (lane == L) {
    u32 above = suf - s;
    bool found = false;
#pragma unroll
    for (int q = kFinePerLane / 4 - 1; q >= 0; --q) {
      const u32 c4[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
      for (int e = 3; e >= 0; --e) {
        if (!found) {
          if (above + c4[e] >= (u32)K) { b = kFinePerLane * L + 4 * q + e; found = true; }
          else above += c4[e];
        }
      }
    }
  }
  return __shfl_sync(0xffffffffu, b, L);
}

// Floor shared between warps working on the same plane (CTA-wide in shared memory, plane-wide in
// global memory).  Unlike the warp-local floor (which may drop equal scores because everything
// it counted has a lower index), a shared floor needs a strict score gap: a pixel is dropped
// only if its logit is below edge(b) - kNear with edge(b) in [kLo, kHi] (Num<DT>), where
// S(x - kNear) < S(x) is verified exhaustively (tests/test_gpu_parity.py).
template <int DT = SDNET_DTYPE_F32>
__device__ __forceinline__ float shared_floor(int fbin, float xscale) {
  if (fbin <= 0) return -CUDART_INF_F;
  const float edge = kBinLo + (float)fbin * (1.0f / kFineScale);
  if (xscale != 1.0f) return edge / xscale;  // pre-activated: keys are strictly monotone in the value
  if (edge < Num<DT>::kLo || edge > Num<DT>::kHi2) return -CUDART_INF_F;
  return edge - (edge <= Num<DT>::kHi ? Num<DT>::kNear : Num<DT>::kNear2);
}

// Where a warp publishes / picks up shared floors.
struct SharedFloors {
  u32* cta_hist;   // shared memory [kFineBins], or nullptr
  int* cta_floor;  // shared memory, or nullptr
  u32* ghist;      // global [kFineBins], or nullptr when the CTA covers the whole plane
  int* gfloor;     // global
};

// Flush the warp's candidate buffer: evaluate the exact score of up to 64 buffered pixels (all
// lanes busy), append (score, index) records to the plane's list with one atomic, feed the
// warp-local and plane-wide histograms and raise the pruning floor.
struct UnitState {
  float floorx;   // input units; a pixel can still matter only if x > floorx
  u32 emitted;    // records this unit has appended so far
  int nbuf;       // records waiting in the shared-memory buffer
};

template <int DT = SDNET_DTYPE_F32>
__device__ __forceinline__ void flush_candidates(UnitState& st, const u64* buf, u32* hist, int* minx,
                                                 const SharedFloors& sf, int* count_ptr, u64* __restrict__ list,
                                                 int cap, int K, int lane, bool pre, float xscale, float satx) {
  const int n = st.nbuf;  // warp-uniform, 1..kBuf
  int base = 0;
  if (lane == 0) base = atomicAdd(count_ptr, n);
  base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
  for (int half = 0; half < kBuf / 32; ++half) {
    const int i = half * 32 + lane;
    if (half * 32 < n) {  // warp-uniform
      const bool valid = i < n;
      const u64 rec = valid ? buf[i] : 0ull;
      const float x = __uint_as_float((u32)(rec >> 32));
      u32 key;
      if (pre) {
        const u32 bits = __float_as_uint(x);
        key = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
      } else {
        key = __float_as_uint(Num<DT>::act(x));
      }
      if (valid) {
        if (base + i < cap) list[base + i] = ((u64)key << 32) | (u32)rec;
        // clamped logit: the score is a monotone function of it, saturation included
        const float xe = fminf(fmaxf(x * xscale, -satx), satx);
        const int bin = logit_bin(xe);
        atomicAdd(&hist[bin], 1u);
        atomicMin(&minx[bin], ord_of(xe));
        const int fb = fine_bin(xe);
        if (sf.cta_hist) atomicAdd(&sf.cta_hist[fb], 1u);
        if (sf.ghist) atomicAdd(&sf.ghist[fb], 1u);
      }
    }
  }
  st.emitted += n;
  st.nbuf = 0;
  __syncwarp();
  // xscale is a power of two, so the division is exact
  if (st.emitted >= (u32)K) st.floorx = fmaxf(st.floorx, local_floor(hist, minx, lane, K) / xscale);
  // publish / refresh the shared floors
  if (sf.cta_hist) {
    const int b = floor_bin_fine<true>(sf.cta_hist, lane, K);
    if (b > 0) {
      if (lane == 0) atomicMax(sf.cta_floor, b);
      st.floorx = fmaxf(st.floorx, shared_floor<DT>(b, xscale));
    }
  }
  if (sf.ghist) {
    const int gb = floor_bin_fine<false>(sf.ghist, lane, K);
    if (gb > 0) {
      if (lane == 0) {
        atomicMax(sf.gfloor, gb);
        if (sf.cta_floor) atomicMax(sf.cta_floor, gb);
      }
      st.floorx = fmaxf(st.floorx, shared_floor<DT>(gb, xscale));
    }
  }
  // a floor at the saturation clamp means "nothing can beat what we have": every x >= satx has
  // the same score as the K recorded ones and a higher index
  if (st.floorx >= satx) st.floorx = CUDART_INF_F;
}

// Which of a lane's four pixels survive NMS, given their window maxima h0..h3 (logit space).
//   x == h            -> certainly survives;
//   x <  h but so close that the two scores may round equal -> settled with the exact score.
// Columns outside the image hold -inf and never pass x > floorx.
template <int R, int DT = SDNET_DTYPE_F32>
__device__ __forceinline__ u32 classify_row(const float4 ctr, float h0, float h1, float h2, float h3, float floorx) {
  constexpr float kNearTie = Num<DT>::kNear, kHiZone = Num<DT>::kHi, kLoZone = Num<DT>::kLo;
  constexpr float kNearTie2 = Num<DT>::kNear2, kHiZone2 = Num<DT>::kHi2;
  u32 cmask = 0, amb = 0;
#define SDNET_CLASSIFY(x, h, j)                                                                        \
  if ((x) > floorx) {                                                                                  \
    if ((x) == (h)) cmask |= 1u << j;                                                                  \
    else if (((x) >= (h) - kNearTie) || ((h) > kHiZone && (x) >= (h) - kNearTie2) ||                    \
             ((h) > kHiZone2 && (x) > kHiZone2 - 1.0f) || ((h) < kLoZone))                              \
      amb |= 1u << j;                                                                                  \
  }
  SDNET_CLASSIFY(ctr.x, h0, 0)
  SDNET_CLASSIFY(ctr.y, h1, 1)
  SDNET_CLASSIFY(ctr.z, h2, 2)
  SDNET_CLASSIFY(ctr.w, h3, 3)
#undef SDNET_CLASSIFY
  // rare: resolve the ambiguous pixels with the exact score function, one per lane per round
  while (__any_sync(0xffffffffu, amb != 0)) {
    const bool has = amb != 0;
    const int jj = has ? __ffs(amb) - 1 : 0;
    const float x = jj == 0 ? ctr.x : (jj == 1 ? ctr.y : (jj == 2 ? ctr.z : ctr.w));
    const float h = jj == 0 ? h0 : (jj == 1 ? h1 : (jj == 2 ? h2 : h3));
    if (has && Num<DT>::act(x) == Num<DT>::act(h)) cmask |= 1u << jj;
    amb &= amb - 1;
  }
  return cmask;
}

// Append the selected pixels of one row as (logit, index) records to the warp's buffer.
// Common case (<= 32 records in the row): positions from three back-to-back ballots on the bits
// of each lane's record count, no branches.  Rows with more (plateaus) go column by column.
template <int DT = SDNET_DTYPE_F32>
__device__ __forceinline__ void append_row(UnitState& st, u32 cmask, const float4 ctr, u32 idx0, u64* buf, u32* hist,
                                           int* minx, const SharedFloors& sf, int* count_ptr,
                                           u64* __restrict__ list, int cap, int K, int lane, bool pre, float xscale,
                                           float satx) {
  const u32 cnt = __popc(cmask);
  const u32 b0 = __ballot_sync(0xffffffffu, cnt & 1u);
  const u32 b1 = __ballot_sync(0xffffffffu, cnt & 2u);
  const u32 b2 = __ballot_sync(0xffffffffu, cnt & 4u);
  const u32 total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
  if (total == 0) return;
  const u32 lt = (1u << lane) - 1u;
  if (total <= 32) {
    if (st.nbuf + (int)total > kBuf) {
      __syncwarp();
      flush_candidates<DT>(st, buf, hist, minx, sf, count_ptr, list, cap, K, lane, pre, xscale, satx);
    }
    int pos = st.nbuf + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
    if (cmask & 1u) buf[pos++] = ((u64)__float_as_uint(ctr.x) << 32) | (idx0 + 0);
    if (cmask & 2u) buf[pos++] = ((u64)__float_as_uint(ctr.y) << 32) | (idx0 + 1);
    if (cmask & 4u) buf[pos++] = ((u64)__float_as_uint(ctr.z) << 32) | (idx0 + 2);
    if (cmask & 8u) buf[pos++] = ((u64)__float_as_uint(ctr.w) << 32) | (idx0 + 3);
    st.nbuf += (int)total;
    return;
  }
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const bool mine = (cmask >> jj) & 1u;
    const u32 m = __ballot_sync(0xffffffffu, mine);
    if (m) {  // warp-uniform
      if (st.nbuf > kBuf - 32) {
        __syncwarp();
        flush_candidates<DT>(st, buf, hist, minx, sf, count_ptr, list, cap, K, lane, pre, xscale, satx);
      }
      if (mine) buf[st.nbuf + __popc(m & lt)] = ((u64)__float_as_uint(comp(ctr, jj)) << 32) | (idx0 + jj);
      st.nbuf += __popc(m);
    }
  }
}

template <bool kAligned, int R, int DT>
__global__ void __launch_bounds__(kThreads, DT == SDNET_DTYPE_F32 ? 4 : 3)
sdnet_peaks_kernel(const __grid_constant__ PeaksParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  unsigned char* wbase = smem_raw + (size_t)warp * kPeaksSmemPerWarp;
  const u32 ring_s = smem_u32(wbase);
  u32* hist = reinterpret_cast<u32*>(wbase + kStages * kPitchB);
  int* minx = reinterpret_cast<int*>(hist + kBins);
  u64* buf = reinterpret_cast<u64*>(minx + kBins);
  const u32 bars_s = smem_u32(buf + kBuf);
  const bool pre = p.pre_activated != 0;
  const float xscale = pre ? kPreScale : 1.0f;
  const float satx = pre ? CUDART_INF_F : kSatX;
  const int C = p.M + p.N;
  const int H = p.H, W = p.W;
  const u32 own_off = (4 + 4 * lane) * 4;                      // this lane's four columns inside a ring row
  const u32 halo_off = (lane == 31 ? 4 + kPanelW : 2) * 4;     // the two columns beyond the panel edge
  typename FeedFor<DT>::type feed;
  feed.init(ring_s, bars_s, lane);

  for (;;) {
    u32 unit = 0;
    if (lane == 0) unit = atomicAdd(p.sched, 1u);
    unit = __shfl_sync(0xffffffffu, unit, 0);
    if (unit >= (u32)p.units) break;
    const int panel = unit % p.panels;
    const int t1 = unit / p.panels;
    const int strip = t1 % p.strips;
    const int plane_id = t1 / p.strips;
    const int b = plane_id / C, c = plane_id % C;
    const bool is_anchor = c < p.M;
    const View4& vw = is_anchor ? p.anchor : p.part;
    const void* plane = static_cast<const typename Num<DT>::In*>(vw.data) + (long long)b * vw.sb +
                        (long long)(is_anchor ? c : c - p.M) * vw.sc;
    const int K = is_anchor ? p.K : p.P;
    const int panel_col0 = panel * kPanelW;
    const int col0 = panel_col0 + lane * 4;
    const int r_begin = strip * p.rows_per_strip;
    const int r_end = min(H, r_begin + p.rows_per_strip);
    const int nrows = r_end - r_begin;
    const int q_last = nrows - 1 + 2 * R;  // ring sequence number of the last row any centre row needs
    u64* __restrict__ list = p.lists + (size_t)plane_id * p.cap;
    int* count_ptr = p.counts + plane_id;
    int* gfloor_ptr = p.gfloor + plane_id;
    SharedFloors sf;
    sf.cta_hist = nullptr; sf.cta_floor = nullptr;
    sf.ghist = p.ghist + (size_t)plane_id * kFineBins;
    sf.gfloor = gfloor_ptr;

    UnitState st;
    st.floorx = shared_floor<DT>(__ldcg(gfloor_ptr), xscale);
    st.emitted = 0;
    st.nbuf = 0;

    __syncwarp();  // everyone is done with the previous unit's ring, histogram and buffer
    *reinterpret_cast<uint4*>(hist + 4 * lane) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<int4*>(minx + 4 * lane) = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff);
    feed.begin_unit(plane, vw.sh, r_begin - R, H, W, panel_col0, q_last, lane);
#pragma unroll
    for (int q = 0; q < kStages - 1; ++q) feed.issue(q, lane);

    for (int t = 0; t < nrows; ++t) {
      // the slot of sequence number t-1 is free: every lane passed a warp-wide vote after reading it
      feed.issue(t + kStages - 1, lane);
      feed.template wait_step<R>();
      __syncwarp();
      const float4 ctr = lds128(feed.slot_addr(t + R) + own_off);
      const float m4 = fmaxf(fmaxf(ctr.x, ctr.y), fmaxf(ctr.z, ctr.w));
      if (__any_sync(0xffffffffu, m4 > st.floorx)) {
        const float floorx = st.floorx;
        u32 cmask = 0;  // bit j: pixel col0+j goes to the candidate buffer
        if (!pre) {
          // vertical (2R+1)-max of own columns and of this lane's halo pair (lanes 0 / 31 only)
          float4 v = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
          float2 hvv = make_float2(-CUDART_INF_F, -CUDART_INF_F);
#pragma unroll
          for (int d = 0; d <= 2 * R; ++d) {
            const u32 so = feed.slot_addr(t + d);
            const float4 o = lds128(so + own_off);
            const float2 ho = lds64(so + halo_off);
            v.x = fmaxf(v.x, o.x); v.y = fmaxf(v.y, o.y); v.z = fmaxf(v.z, o.z); v.w = fmaxf(v.w, o.w);
            hvv.x = fmaxf(hvv.x, ho.x); hvv.y = fmaxf(hvv.y, ho.y);
          }
          float L2 = __shfl_up_sync(0xffffffffu, v.z, 1);
          float L3 = __shfl_up_sync(0xffffffffu, v.w, 1);
          float R0 = __shfl_down_sync(0xffffffffu, v.x, 1);
          float R1 = __shfl_down_sync(0xffffffffu, v.y, 1);
          if (lane == 0) { L2 = hvv.x; L3 = hvv.y; }
          if (lane == 31) { R0 = hvv.x; R1 = hvv.y; }
          float h0, h1, h2, h3;
          if (R == 2) {
            const float m12 = fmaxf(v.y, v.z);
            h0 = max3(fmaxf(L2, L3), v.x, m12);
            h1 = max3(fmaxf(L3, v.x), m12, v.w);
            h2 = max3(fmaxf(v.x, R0), m12, v.w);
            h3 = max3(fmaxf(R0, R1), m12, v.w);
          } else {
            h0 = max3(L3, v.x, v.y);
            h1 = max3(v.x, v.y, v.z);
            h2 = max3(v.y, v.z, v.w);
            h3 = max3(v.z, v.w, R0);
          }
          cmask = classify_row<R, DT>(ctr, h0, h1, h2, h3, floorx);
        } else {
          // pre-activated maps (CoreMLDecoder): every pixel above the floor is a candidate
          if (ctr.x > floorx) cmask |= 1u;
          if (ctr.y > floorx) cmask |= 2u;
          if (ctr.z > floorx) cmask |= 4u;
          if (ctr.w > floorx) cmask |= 8u;
        }
        append_row<DT>(st, cmask, ctr, (u32)((r_begin + t) * W + col0), buf, hist, minx, sf, count_ptr, list,
                   p.cap, K, lane, pre, xscale, satx);
      }
      if ((t & 7) == 7) {
        // every 8 rows: pick up the plane-wide floor other warps may have raised
        st.floorx = fmaxf(st.floorx, shared_floor<DT>(__ldcg(gfloor_ptr), xscale));
      }
    }
    feed.end_unit();
    if (st.nbuf) {
      __syncwarp();
      flush_candidates<DT>(st, buf, hist, minx, sf, count_ptr, list, p.cap, K, lane, pre, xscale, satx);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// peaks kernel, warp-specialised form (the fast path; needs 16 B-aligned rows and W <= 896)
//
// CTA = NC consumer warps + 1 producer warp, working on one (plane, row strip) unit at a time.
// The producer streams whole image rows into a shared-memory ring with 1-D bulk copies (TMA
// unit, SASS UBLKCP): four rows per mbarrier ("group"), and -- when the plane is dense in
// memory -- ONE copy of 4*W*4 contiguous bytes per group.  Ring rows are dense (W floats), so
// horizontal neighbours, also across warps, are plain shared-memory reads; the two image-edge
// cases are patched with -inf by the few lanes that touch them.  Consumer warp w owns columns
// [128 w, 128 w + 128).  Per group of four output rows it waits on one barrier, reads its four
// centre rows (4 x LDS.128), takes the max of the 16 values and votes "does anything here beat
// the pruning floor?"; only then does it look at single rows.
// ---------------------------------------------------------------------------------------------
constexpr int kGroupRows = 4;
constexpr int kMaxConsumers = 7;

struct CtaGeom {
  int nc;       // consumer warps
  int rpb;      // ring row pitch in bytes (= W * 4)
  int ng;       // ring groups
  int smem;     // dynamic shared memory per CTA
};

inline CtaGeom cta_geometry(int W, int ng) {
  CtaGeom g;
  g.nc = (W + kPanelW - 1) / kPanelW;
  g.rpb = W * 4;
  g.ng = ng;
  // ring, 512 B of slack (lanes past the last column read harmlessly beyond their row),
  // barriers, per-consumer pruning state
  g.smem = ng * kGroupRows * g.rpb + 512 + 2 * ng * 8 + 2 * kFineBins * 4 + 64 + g.nc * (kBins * 8 + kBuf * 8);
  return g;
}

// Window maxima of a lane's four pixels for output row t, straight from the ring.  Ring rows
// are contiguous in shared memory, so ring row q of a unit whose first group has running number
// n lives at row (4 n + q) mod (4 NG): `rowbase` = 4 n, `rowmask` = 4 NG - 1.
template <int R, u32 kRows>
__device__ __forceinline__ void window_max(u32 ring_own, u32 rowbase, int rpb, int t, bool left_ok,
                                           bool right_ok, float& h0, float& h1, float& h2, float& h3) {
  const float ninf = -CUDART_INF_F;
  float v0 = ninf, v1 = ninf, v2 = ninf, v3 = ninf, v4 = ninf, v5 = ninf, v6 = ninf, v7 = ninf;  // cols c-2 .. c+5
  const u32 loff = left_ok ? 8u : 0u;  // column 0 has no left neighbours (and nothing mapped before the ring)
#pragma unroll
  for (int d = 0; d <= 2 * R; ++d) {
    const u32 a = ring_own + ((rowbase + (u32)(t + d)) % kRows) * rpb;  // kRows is a constant: AND when a power of two
    const float4 o = lds128(a);
    const float2 l = lds64(a - loff);
    const float2 r = lds64(a + 16);
    v0 = fmaxf(v0, l.x); v1 = fmaxf(v1, l.y);
    v2 = fmaxf(v2, o.x); v3 = fmaxf(v3, o.y); v4 = fmaxf(v4, o.z); v5 = fmaxf(v5, o.w);
    v6 = fmaxf(v6, r.x); v7 = fmaxf(v7, r.y);
  }
  if (!left_ok) { v0 = ninf; v1 = ninf; }    // beyond the left image edge: max_pool2d's -inf padding
  if (!right_ok) { v6 = ninf; v7 = ninf; }   // beyond the right image edge
  if (R == 2) {
    const float m34 = fmaxf(v3, v4);
    h0 = max3(fmaxf(v0, v1), v2, m34);
    h1 = max3(fmaxf(v1, v2), m34, v5);
    h2 = max3(fmaxf(v2, v6), m34, v5);
    h3 = max3(fmaxf(v6, v7), m34, v5);
  } else {
    h0 = max3(v1, v2, v3);
    h1 = max3(v2, v3, v4);
    h2 = max3(v3, v4, v5);
    h3 = max3(v4, v5, v6);
  }
}

template <int R, int NG>
__global__ void __launch_bounds__((kMaxConsumers + 1) * 32, 3) sdnet_peaks_cta_kernel(const __grid_constant__ PeaksParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  static_assert((NG & (NG - 1)) == 0, "ring groups must be a power of two");
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nc = (int)(blockDim.x >> 5) - 1;
  const int C = p.M + p.N;
  const int H = p.H, W = p.W;
  const int rpb = W * 4;
  const u32 ring_s = smem_u32(smem_raw);
  const u32 full_s = ring_s + NG * kGroupRows * rpb + 512;
  const u32 empty_s = full_s + NG * 8;
  // CTA-wide pruning state, double-buffered by unit parity: [2][kFineBins] histogram, then
  // floor bin [2] and finished-consumer count [2]
  u32* cta_hist = reinterpret_cast<u32*>(smem_raw + NG * kGroupRows * rpb + 512 + 2 * NG * 8);
  int* cta_floor = reinterpret_cast<int*>(cta_hist + 2 * kFineBins);
  volatile int* cta_done = cta_floor + 2;
  unsigned char* wstate = reinterpret_cast<unsigned char*>(cta_hist + 2 * kFineBins) + 64;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NG; ++i) {
      mbar_init(full_s + 8 * i, 1);
      mbar_init(empty_s + 8 * i, (u32)nc);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 2 * kFineBins + 16; i += blockDim.x) cta_hist[i] = 0;  // histograms + control words
  __syncthreads();

  if (warp == nc) {
    // ================================ producer warp ================================
    const u32 row_bytes = (u32)rpb;
    u32 n = 0;  // running group number: slot n % NG, phase (n / NG) & 1
    int uo = 0;  // ordinal of the unit within this CTA
    for (u32 unit = blockIdx.x; unit < (u32)p.units; unit += gridDim.x, ++uo) {
      if (uo >= 2) {
        // recycle the CTA-wide histogram of unit uo-2: wait until all its consumers are done
        // (they are, unless units are only a few rows long), then clear it.  Consumers see the
        // cleared state through the release/acquire of this unit's first full barrier.
        const int slot = uo & 1;
        const int want = nc * (uo >> 1);
        if (lane == 0) while (cta_done[slot] < want) __nanosleep(64);
        __syncwarp();
        for (int i = lane; i < kFineBins; i += 32) cta_hist[slot * kFineBins + i] = 0;
        if (lane == 0) cta_floor[slot] = 0;
        __syncwarp();
      }
      const int strip = unit % p.strips;
      const int plane_id = unit / p.strips;
      const int b = plane_id / C, c = plane_id % C;
      const bool is_anchor = c < p.M;
      const View4& vw = is_anchor ? p.anchor : p.part;
      const float* plane = static_cast<const float*>(vw.data) + (long long)b * vw.sb + (long long)(is_anchor ? c : c - p.M) * vw.sc;
      const int r_begin = strip * p.rows_per_strip;
      const int r_end = min(H, r_begin + p.rows_per_strip);
      const int q_count = r_end - r_begin + 2 * R;  // ring rows of this unit: image rows r_begin-R .. r_end-1+R
      const int q_lo = max(0, R - r_begin);          // first ring row that lies inside the image
      const int q_hi = min(q_count, H - (r_begin - R));  // one past the last ring row inside the image
      const int groups = (q_count + kGroupRows - 1) / kGroupRows;
      const long long pitch = vw.sh * 4;
      const bool dense = pitch == (long long)row_bytes;
      const char* src0 = reinterpret_cast<const char*>(plane) + (long long)(r_begin - R) * pitch;
      // Optional L2 prefetch ahead of the ring (SDNET_L2_PREFETCH_GROUPS): DRAM latency is then
      // absorbed by L2 rather than by the small shared-memory ring.
      const int pf = p.l2_prefetch_groups;
      if (pf > 0 && dense && lane == 0) {
        const int v1p = min(min(pf * kGroupRows, q_count), q_hi);
        if (v1p > q_lo) bulk_prefetch_l2(src0 + (long long)q_lo * pitch, (u32)(v1p - q_lo) * row_bytes);
      }
      for (int j = 0; j < groups; ++j, ++n) {
        const u32 slot = n & (NG - 1);
        const int q0 = j * kGroupRows, q1 = min(q0 + kGroupRows, q_count);
        if (pf > 0 && dense && lane == 0) {
          const int pq0 = max((j + pf) * kGroupRows, q_lo), pq1 = min(min((j + pf + 1) * kGroupRows, q_count), q_hi);
          if (pq1 > pq0) bulk_prefetch_l2(src0 + (long long)pq0 * pitch, (u32)(pq1 - pq0) * row_bytes);
        }
        const int v0 = max(q0, q_lo), v1 = min(q1, q_hi);  // [v0, v1) = rows of this group inside the image
        mbar_wait(empty_s + 8 * slot, ((n / NG) & 1u) ^ 1u);  // consumers are done with this slot
        if (v0 > q0 || v1 < q1) {  // rows above / below the image: -inf (max_pool2d's padding)
          for (int q = q0; q < q1; ++q)
            if (q < v0 || q >= v1)
              for (int k = lane; k < rpb / 16; k += 32) sts128(ring_s + (slot * kGroupRows + (q - q0)) * rpb + 16 * k, -CUDART_INF_F);
          fence_proxy_async();
          __syncwarp();
        }
        if (lane == 0) {
          const u32 bar = full_s + 8 * slot;
          if (v1 > v0) {
            mbar_arrive_expect_tx(bar, (u32)(v1 - v0) * row_bytes);
            if (dense) {
              bulk_g2s(ring_s + (slot * kGroupRows + (v0 - q0)) * rpb, src0 + (long long)v0 * pitch,
                       (u32)(v1 - v0) * row_bytes, bar);
            } else {
              for (int q = v0; q < v1; ++q)
                bulk_g2s(ring_s + (slot * kGroupRows + (q - q0)) * rpb, src0 + (long long)q * pitch, row_bytes, bar);
            }
          } else {
            mbar_arrive(bar);
          }
        }
        __syncwarp();
      }
    }
    return;
  }

  // ================================== consumer warps ==================================
  u32* hist = reinterpret_cast<u32*>(wstate + (size_t)warp * (kBins * 8 + kBuf * 8));
  int* minx = reinterpret_cast<int*>(hist + kBins);
  u64* buf = reinterpret_cast<u64*>(minx + kBins);
  const bool pre = p.pre_activated != 0;
  const float xscale = pre ? kPreScale : 1.0f;
  const float satx = pre ? CUDART_INF_F : kSatX;
  const int col0 = kPanelW * warp + 4 * lane;
  const bool lane_ok = col0 < W;          // W % 4 == 0: a lane's four columns are all inside or all outside
  const bool left_ok = col0 > 0;
  const bool right_ok = col0 + 4 < W;
  const u32 ring_own = ring_s + (u32)col0 * 4;
  const float ninf = -CUDART_INF_F;
  u32 n = 0;
  int uo = 0;
  for (u32 unit = blockIdx.x; unit < (u32)p.units; unit += gridDim.x, ++uo) {
    const int strip = unit % p.strips;
    const int plane_id = unit / p.strips;
    const int c = plane_id % C;
    const int K = c < p.M ? p.K : p.P;
    const int r_begin = strip * p.rows_per_strip;
    const int r_end = min(H, r_begin + p.rows_per_strip);
    const int nrows = r_end - r_begin;
    const int groups = (nrows + 2 * R + kGroupRows - 1) / kGroupRows;
    const int groups_out = (nrows + kGroupRows - 1) / kGroupRows;
    u64* __restrict__ list = p.lists + (size_t)plane_id * p.cap;
    int* count_ptr = p.counts + plane_id;
    int* gfloor_ptr = p.gfloor + plane_id;
    SharedFloors sf;
    sf.cta_hist = cta_hist + (uo & 1) * kFineBins;
    sf.cta_floor = cta_floor + (uo & 1);
    // with one strip per plane this CTA sees the whole plane: no plane-wide exchange needed
    sf.ghist = p.strips > 1 ? p.ghist + (size_t)plane_id * kFineBins : nullptr;
    sf.gfloor = gfloor_ptr;
    const volatile int* cta_floor_v = sf.cta_floor;

    UnitState st;
    st.floorx = p.strips > 1 ? shared_floor(__ldcg(gfloor_ptr), xscale) : -CUDART_INF_F;
    st.emitted = 0;
    st.nbuf = 0;
    __syncwarp();
    *reinterpret_cast<uint4*>(hist + 4 * lane) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<int4*>(minx + 4 * lane) = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff);
    __syncwarp();

    const u32 rowbase = n * kGroupRows;
    constexpr u32 kRowMask = NG * kGroupRows - 1;
    int gfloor_seen = 0;
    mbar_wait(full_s + 8 * (n & (NG - 1)), (n / NG) & 1u);
    for (int g = 0; g < groups_out; ++g) {
      const u32 n0 = n + g, n1 = n0 + 1;
      if (g + 1 < groups) mbar_wait(full_s + 8 * (n1 & (NG - 1)), (n1 / NG) & 1u);
      // the floor the other warps of this CTA have reached (shared memory, always fresh) ...
      st.floorx = fmaxf(st.floorx, shared_floor(*cta_floor_v, xscale));
      if (p.strips > 1 && (g & 7) == 0) {
        // ... and, every 32 rows, the one other CTAs working on this plane published (global
        // memory; applied one period late so the load latency is never exposed)
        st.floorx = fmaxf(st.floorx, shared_floor(gfloor_seen, xscale));
        gfloor_seen = __ldcg(gfloor_ptr);
      }
      const int t0 = g * kGroupRows;
      // centre row of output row t0+i is ring row t0+i+R
      const u32 a0 = ring_own + ((rowbase + (u32)(t0 + R)) & kRowMask) * rpb;
      const u32 a1 = ring_own + ((rowbase + (u32)(t0 + R + 1)) & kRowMask) * rpb;
      const u32 a2 = ring_own + ((rowbase + (u32)(t0 + R + 2)) & kRowMask) * rpb;
      const u32 a3 = ring_own + ((rowbase + (u32)(t0 + R + 3)) & kRowMask) * rpb;
      const int rows_here = min(kGroupRows, nrows - t0);
      const float4 c0 = lds128(a0);
      const float4 c1 = lds128(a1);   // rows past the strip hold stale data: masked out just below
      const float4 c2 = lds128(a2);
      const float4 c3 = lds128(a3);
      float m0 = fmaxf(fmaxf(c0.x, c0.y), fmaxf(c0.z, c0.w));
      float m1 = fmaxf(fmaxf(c1.x, c1.y), fmaxf(c1.z, c1.w));
      float m2 = fmaxf(fmaxf(c2.x, c2.y), fmaxf(c2.z, c2.w));
      float m3 = fmaxf(fmaxf(c3.x, c3.y), fmaxf(c3.z, c3.w));
      if (rows_here < kGroupRows) {  // last, partial group of the strip (warp-uniform)
        if (rows_here < 2) m1 = ninf;
        if (rows_here < 3) m2 = ninf;
        m3 = ninf;
      }
      const float m = lane_ok ? fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) : ninf;
      if (__any_sync(0xffffffffu, m > st.floorx)) {
        // something in these four rows beats the floor: visit only the rows that do (kept as a
        // real loop so the row code exists once -- it is large and instruction-cache bound)
        u32 rowmask4 = (__any_sync(0xffffffffu, lane_ok && m0 > st.floorx) ? 1u : 0u) |
                       (__any_sync(0xffffffffu, lane_ok && m1 > st.floorx) ? 2u : 0u) |
                       (__any_sync(0xffffffffu, lane_ok && m2 > st.floorx) ? 4u : 0u) |
                       (__any_sync(0xffffffffu, lane_ok && m3 > st.floorx) ? 8u : 0u);
#pragma unroll 1
        while (rowmask4) {
          const int i = __ffs(rowmask4) - 1;
          rowmask4 &= rowmask4 - 1;
          const int t = t0 + i;
          float4 ctr = lds128(ring_own + ((rowbase + (u32)(t + R)) & kRowMask) * rpb);
          if (!lane_ok) ctr = make_float4(ninf, ninf, ninf, ninf);
          const float floorx = st.floorx;
          const float mi = fmaxf(fmaxf(ctr.x, ctr.y), fmaxf(ctr.z, ctr.w));
          if (!__any_sync(0xffffffffu, mi > floorx)) continue;  // an earlier row of the group raised the floor
          u32 cmask = 0;
          if (!pre) {
            float h0, h1, h2, h3;
            window_max<R, NG * kGroupRows>(ring_own, rowbase, rpb, t, left_ok, right_ok, h0, h1, h2, h3);
            cmask = classify_row<R>(ctr, h0, h1, h2, h3, floorx);
          } else {
            if (ctr.x > floorx) cmask |= 1u;
            if (ctr.y > floorx) cmask |= 2u;
            if (ctr.z > floorx) cmask |= 4u;
            if (ctr.w > floorx) cmask |= 8u;
          }
          append_row(st, cmask, ctr, (u32)((r_begin + t) * W + col0), buf, hist, minx, sf, count_ptr,
                     list, p.cap, K, lane, pre, xscale, satx);
        }
      }
      // every lane's reads of group n0 are consumed (the votes above): hand the slot back
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_s + 8 * (n0 & (NG - 1)));
      // while the CTA has no floor yet, publish early and often; later only in batches
      if (st.nbuf >= 16 || (st.nbuf > 0 && *cta_floor_v == 0)) {
        __syncwarp();
        flush_candidates(st, buf, hist, minx, sf, count_ptr, list, p.cap, K, lane, pre, xscale, satx);
      }
    }
    if (groups > groups_out) {  // a trailing group that only held the bottom halo rows
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_s + 8 * ((n + groups_out) & (NG - 1)));
    }
    n += (u32)groups;
    if (st.nbuf) {
      __syncwarp();
      flush_candidates(st, buf, hist, minx, sf, count_ptr, list, p.cap, K, lane, pre, xscale, satx);
    }
    __syncwarp();
    if (lane == 0) atomicAdd(const_cast<int*>(cta_done) + (uo & 1), 1);  // this warp is done with the unit's shared state
  }
}

// ---------------------------------------------------------------------------------------------
// peaks kernel, TMA-tile form (the default fast path; needs 16 B-aligned rows)
//
// Every warp is an autonomous pipeline over one (plane, row strip, 128-column panel) unit.
// An elected lane pulls 4-row x 136-column tiles (the panel plus four columns either side)
// into the warp's private shared-memory ring with 2-D tensor-map bulk copies (TMA, SASS
// UTMALDG), completion on one mbarrier per ring slot.  The tensor map is encoded with NaN
// out-of-bounds fill: fmaxf ignores a NaN operand and every ordered comparison with NaN is
// false, so out-of-image rows and columns behave exactly like max_pool2d's -inf padding with
// no edge code at all.  Per group of four output rows the warp waits on one barrier, reads its
// four centre rows (4 x LDS.128), takes the max of the 16 values and votes "does anything here
// beat the pruning floor?"; only then does it look at single rows.  No warp ever waits for
// another warp.
// ---------------------------------------------------------------------------------------------
constexpr int kTileCols = kPanelW + 8;
constexpr int kTilePitchB = kTileCols * 4;                   // 544 B per ring row
constexpr int kTileBytes = kGroupRows * kTilePitchB;         // 2176 B per TMA tile (17 x 128 B)
constexpr int kTileWarps = 4;
constexpr int kTileNG = 4;   // ring slots (tiles) per warp: 16 rows, two or three tiles in flight, 5 CTAs/SM
constexpr int kWork = 128;   // per-warp work list: one byte per (row of the group, lane) whose 16-byte word holds a pixel above the floor
// S = rows per TMA row (see the kernel).  Under S = 2 a tile arrives as two 2-row boxes and a TMA
// destination must be 128-byte aligned: the second box sits at +1152 and a slot takes 2304 bytes.
__host__ __device__ constexpr int tile_slot_bytes(int S) { return S == 1 ? kTileBytes : 2304; }
__host__ __device__ constexpr int tile_smem_per_warp(int S) {
  return ((kTileNG * tile_slot_bytes(S) + 32 + kBins * 8 + kBuf * 8 + kWork) + 127) / 128 * 128;
}
__host__ __device__ constexpr int tile_smem(int S) { return kTileWarps * tile_smem_per_warp(S); }
constexpr int kOddBoxOff = 1152;  // S = 2: offset of the odd rows' box inside a slot
constexpr int kOddShiftB = 8;     // S = 2: a box must start on a 16-byte boundary of global memory and odd rows start 8 bytes
                                  // off one, so their box starts 4 columns early and their pixels sit 8 bytes further right

// Element geometry of the tile kernel.  A lane owns one 16-byte word per row: 4 fp32 or 8 fp16/bf16
// pixels, so a warp's panel is 128 or 256 columns and a ring row is 544 bytes either way.
template <int DT>
struct TileGeom {
  static constexpr int kPx = DT == SDNET_DTYPE_F32 ? 4 : 8;  // pixels per lane per row = halo columns each side
  static constexpr int kEsz = 16 / kPx;                       // bytes per element
  static constexpr int kPanel = 32 * kPx;                     // columns per warp
  static constexpr int kCols = kPanel + 2 * kPx;              // columns per tile row
};
static_assert(TileGeom<SDNET_DTYPE_F32>::kCols * 4 == kTilePitchB && TileGeom<SDNET_DTYPE_F16>::kCols * 2 == kTilePitchB, "ring row pitch");

__device__ __forceinline__ void tma_tile_4d(u32 dst, const CUtensorMap* map, int x, int y, int c, int b, u32 bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(c), "r"(b), "r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds64x2(u32 addr) {  // 16 bytes from an 8-byte-aligned address
  uint4 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.z), "=r"(v.w) : "r"(addr + 8));
  return v;
}
__device__ __forceinline__ uint4 lds128u(u32 addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// max of the 4 | 8 elements of one 16-byte word / of four words, as float; NaN elements (the TMA
// out-of-bounds fill) are ignored by fmaxf and by max.f16x2 / max.bf16x2 alike
template <int DT>
struct TileMax;
template <>
struct TileMax<SDNET_DTYPE_F32> {
  static __device__ __forceinline__ float word(const uint4& a) {
    return fmaxf(fmaxf(__uint_as_float(a.x), __uint_as_float(a.y)), fmaxf(__uint_as_float(a.z), __uint_as_float(a.w)));
  }
  static __device__ __forceinline__ float group(const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
    return fmaxf(fmaxf(word(a), word(b)), fmaxf(word(c), word(d)));
  }
  static __device__ __forceinline__ float elem(u32 addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
  }
  // one element in its storage format (here: the float's bits), max in that format, back to float
  static __device__ __forceinline__ u32 raw(u32 addr) { return __float_as_uint(elem(addr)); }
  static __device__ __forceinline__ u32 rmax(u32 a, u32 b) { return __float_as_uint(fmaxf(__uint_as_float(a), __uint_as_float(b))); }
  static __device__ __forceinline__ float rfloat(u32 r) { return __uint_as_float(r); }
};
template <>
struct TileMax<SDNET_DTYPE_F16> {
  static __device__ __forceinline__ __half2 h2(u32 v) { return *reinterpret_cast<const __half2*>(&v); }
  static __device__ __forceinline__ __half2 word2(const uint4& a) { return __hmax2(__hmax2(h2(a.x), h2(a.y)), __hmax2(h2(a.z), h2(a.w))); }
  static __device__ __forceinline__ float fold(__half2 m) { return __half2float(__hmax(__low2half(m), __high2half(m))); }
  static __device__ __forceinline__ float word(const uint4& a) { return fold(word2(a)); }
  static __device__ __forceinline__ float group(const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
    return fold(__hmax2(__hmax2(word2(a), word2(b)), __hmax2(word2(c), word2(d))));
  }
  static __device__ __forceinline__ float elem(u32 addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return __half2float(__ushort_as_half(v));
  }
  // storage format: the 16 bits in the low half of a register (high half +0); max.f16x2 ignores NaN
  static __device__ __forceinline__ u32 raw(u32 addr) {
    u32 v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
  }
  static __device__ __forceinline__ u32 rmax(u32 a, u32 b) {
    const __half2 m = __hmax2(h2(a), h2(b));
    return *reinterpret_cast<const u32*>(&m);
  }
  static __device__ __forceinline__ float rfloat(u32 r) { return __half2float(__ushort_as_half((unsigned short)r)); }
};
template <>
struct TileMax<SDNET_DTYPE_BF16> {
  static __device__ __forceinline__ __nv_bfloat162 h2(u32 v) { return *reinterpret_cast<const __nv_bfloat162*>(&v); }
  static __device__ __forceinline__ __nv_bfloat162 word2(const uint4& a) { return __hmax2(__hmax2(h2(a.x), h2(a.y)), __hmax2(h2(a.z), h2(a.w))); }
  static __device__ __forceinline__ float fold(__nv_bfloat162 m) { return __bfloat162float(__hmax(__low2bfloat16(m), __high2bfloat16(m))); }
  static __device__ __forceinline__ float word(const uint4& a) { return fold(word2(a)); }
  static __device__ __forceinline__ float group(const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
    return fold(__hmax2(__hmax2(word2(a), word2(b)), __hmax2(word2(c), word2(d))));
  }
  static __device__ __forceinline__ float elem(u32 addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return __uint_as_float((u32)v << 16);
  }
  static __device__ __forceinline__ u32 raw(u32 addr) {
    u32 v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
  }
  static __device__ __forceinline__ u32 rmax(u32 a, u32 b) {
    const __nv_bfloat162 m = __hmax2(h2(a), h2(b));
    return *reinterpret_cast<const u32*>(&m);
  }
  static __device__ __forceinline__ float rfloat(u32 r) { return __uint_as_float(r << 16); }
};

// Byte offset of ring row rr (0..15) inside the ring.  S = 1: rows in order.  S = 2 (rows loaded as
// even/odd pairs, see the kernel): a slot holds rows 0, 2 at +0, +544 and rows 1, 3 at +1152, +1696,
// the odd rows shifted right by kOddShiftB bytes.
template <int S>
__device__ __forceinline__ u32 ring_row_off(u32 rr) {
  if (S == 1) return rr * kTilePitchB;
  return (rr >> 2) * tile_slot_bytes(2) + (rr & 1u) * (kOddBoxOff + kOddShiftB) + ((rr >> 1) & 1u) * kTilePitchB;
}

// Settle the work list of one 4-row group.  Entry e = (row in group << 5) | lane names one 16-byte
// word of centre pixels holding at least one pixel above the floor; kPx consecutive lanes take the
// pixels of an entry, so records leave in (row, column) = index order.  A lane whose pixel beats the
// floor reads the pixel's (2R+1)^2 window from the ring with scalar loads (columns and rows outside
// the image hold NaN or -inf and never win a max), classifies it like classify_row and appends a
// (logit, index) record to the warp's candidate buffer.
// `row0` = ring row of the window's first row for group row 0 (the centre is R rows further).
template <int R, int DT, int S>
__device__ __forceinline__ void settle_entries(UnitState& st, const unsigned char* work, int nent, u32 ring_s, u32 row0,
                                               float floorx, u32 idx0, int W, bool pre, u64* buf, u32* hist, int* minx,
                                               const SharedFloors& sf, int* count_ptr, u64* __restrict__ list, int cap,
                                               int K, int lane, float xscale, float satx) {
  constexpr float kNearTie = Num<DT>::kNear, kHiZone = Num<DT>::kHi, kLoZone = Num<DT>::kLo, kNearTie2 = Num<DT>::kNear2,
                  kHiZone2 = Num<DT>::kHi2;
  constexpr u32 kRowMask = kTileNG * kGroupRows - 1;
  constexpr int kPx = TileGeom<DT>::kPx, kEsz = TileGeom<DT>::kEsz;
  const int nslots = kPx * nent;
  for (int base = 0; base < nslots; base += 32) {  // warp-uniform
    const int slot = base + lane;
    const u32 e = slot < nslots ? work[slot / kPx] : 0u;
    const u32 i = e >> 5, colp = kPx * (e & 31u) + (u32)(slot % kPx);
    const u32 col_addr = ring_s + (kPx + colp - R) * kEsz;  // first column of the window
    const float x = TileMax<DT>::elem(col_addr + R * kEsz + ring_row_off<S>((row0 + i + R) & kRowMask));
    bool keep = slot < nslots && x > floorx;
    if (keep && !pre) {
      // window max in the storage format (no conversions for fp16/bf16), one accumulator per window row
      u32 hr[2 * R + 1];
#pragma unroll
      for (int d = 0; d <= 2 * R; ++d) {
        const u32 a = col_addr + ring_row_off<S>((row0 + i + d) & kRowMask);
        u32 v[2 * R + 1];
#pragma unroll
        for (int q = 0; q <= 2 * R; ++q) v[q] = TileMax<DT>::raw(a + kEsz * q);
        hr[d] = v[0];
#pragma unroll
        for (int q = 1; q <= 2 * R; ++q) hr[d] = TileMax<DT>::rmax(hr[d], v[q]);
      }
#pragma unroll
      for (int d = 1; d <= 2 * R; ++d) hr[0] = TileMax<DT>::rmax(hr[0], hr[d]);
      const float h = fmaxf(x, TileMax<DT>::rfloat(hr[0]));  // an all-NaN window cannot happen: the centre is in it
      if (x != h) {
        const bool amb = (x >= h - kNearTie) || (h > kHiZone && x >= h - kNearTie2) ||
                         (h > kHiZone2 && x > kHiZone2 - 1.0f) || (h < kLoZone);  // same zones as classify_row
        keep = amb && Num<DT>::act(x) == Num<DT>::act(h);  // rare
      }
    }
    const u32 m = __ballot_sync(0xffffffffu, keep);
    if (m) {  // warp-uniform
      if (st.nbuf + __popc(m) > kBuf) {
        __syncwarp();
        flush_candidates<DT>(st, buf, hist, minx, sf, count_ptr, list, cap, K, lane, pre, xscale, satx);
      }
      if (keep) buf[st.nbuf + __popc(m & ((1u << lane) - 1u))] = ((u64)__float_as_uint(x) << 32) | (idx0 + i * (u32)W + colp);
      st.nbuf += __popc(m);
    }
  }
}

// S = 2 only: a tile that touches the left or right image edge has read across a row boundary (see
// the kernel): overwrite what is not this row's data with -inf.  `slot_s` = the tile's ring slot.
template <int DT>
__device__ __forceinline__ void tile_fix_edges(u32 slot_s, int x0, int W, int lane) {
  constexpr int kPx = TileGeom<DT>::kPx;
  constexpr u32 kNinf2 = DT == SDNET_DTYPE_F16 ? 0xFC00FC00u : (DT == SDNET_DTYPE_BF16 ? 0xFF80FF80u : 0xFF800000u);
  if (x0 < 0) {  // odd rows (the second box): the 12 columns left of column 0 hold the end of the row above
    if (lane < 2) {
      const u32 a = slot_s + kOddBoxOff + lane * kTilePitchB;
      asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(kNinf2) : "memory");
      asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(a + 16), "r"(kNinf2) : "memory");
    }
  }
  if (x0 + TileGeom<DT>::kCols > W) {  // even rows (the first box): columns >= W hold the start of the row below
    // word w of a ring row covers columns x0 + kPx w ..; under S = 2 W is a multiple of kPx/2, not of kPx
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int w = k == 0 ? lane + 1 : 33;  // lane's own centre word; lane 31 also takes the right halo word
      if (k == 1 && lane != 31) break;
      const int first = x0 + kPx * w;
      const u32 a = slot_s + 16 * w;
      if (first >= W) {
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(kNinf2) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a + kTilePitchB), "r"(kNinf2) : "memory");
      } else if (first + kPx / 2 >= W) {
        asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(a + 8), "r"(kNinf2) : "memory");
        asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(a + 8 + kTilePitchB), "r"(kNinf2) : "memory");
      }
    }
  }
}

// S = rows per TMA row: 1 when the row pitch is a multiple of 16 bytes.  S = 2 serves fp16/bf16 maps
// whose pitch is an odd multiple of 8 bytes (W = 612): the tensor map then describes PAIRS of image
// rows as one row of pitch + W elements, a tile is two 2-row boxes -- the even rows at x, the odd rows
// at pitch + x - 4 (a box has to start on a 16-byte boundary, measured: anything else is an illegal
// instruction) -- and lands in its slot as rows 0, 2 | 1, 3 with the odd rows 8 bytes further right.
// At the image edges such a box reads across the row boundary; tile_fix_edges repairs that after the wait.
template <int R, int DT, int S>
__global__ void __launch_bounds__(kTileWarps * 32, 5)
sdnet_peaks_tile_kernel(const __grid_constant__ PeaksParams p, const __grid_constant__ CUtensorMap tm_anchor,
                        const __grid_constant__ CUtensorMap tm_part) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NG = kTileNG;
  constexpr u32 kRowMask = NG * kGroupRows - 1;
  constexpr int kPx = TileGeom<DT>::kPx, kPanel = TileGeom<DT>::kPanel;
  static_assert((NG & (NG - 1)) == 0, "slot and parity of a tile come from its running number by mask and shift");
  static_assert(S == 1 || (R == 2 && DT != SDNET_DTYPE_F32), "row pairs: tiles must start on an even row");
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr u32 kSlotB = tile_slot_bytes(S);
  unsigned char* wbase = smem_raw + (size_t)warp * tile_smem_per_warp(S);
  const u32 ring_s = smem_u32(wbase);
  const u32 bars_s = ring_s + NG * kSlotB;
  u32* hist = reinterpret_cast<u32*>(wbase + NG * kSlotB + 32);
  int* minx = reinterpret_cast<int*>(hist + kBins);
  u64* buf = reinterpret_cast<u64*>(minx + kBins);
  unsigned char* work = reinterpret_cast<unsigned char*>(buf + kBuf);
  const bool pre = p.pre_activated != 0;
  const float xscale = pre ? kPreScale : 1.0f;
  const float satx = pre ? CUDART_INF_F : kSatX;
  const int C = p.M + p.N;
  const int H = p.H, W = p.W;
  const u32 ring_own = ring_s + (u32)(16 + 16 * lane);  // this lane's word inside a ring row
  const u32 lt = (1u << lane) - 1u;

  if (lane == 0) {
    for (int i = 0; i < NG; ++i) mbar_init(bars_s + 8 * i, 1);
    mbar_fence_init();
  }
  __syncwarp();

  // Tiles are numbered in one running sequence over all units this warp processes: tile number n
  // lives in ring slot n % NG (ring rows 4 (n % NG) ..) and is the (n / NG)-th use of that slot, so
  // the slot's mbarrier is waited with parity (n / NG) & 1.  Every issued tile is waited exactly once.
  u32 tile_n = 0;
  for (;;) {
    u32 unit = 0;
    if (lane == 0) unit = atomicAdd(p.sched, 1u);
    unit = __shfl_sync(0xffffffffu, unit, 0);
    if (unit >= (u32)p.units) break;
    int panel, plane_id, r_begin, r_end;
    if (unit < (u32)p.tier1_units) {
      panel = unit % p.panels;
      plane_id = unit / p.panels;
      r_begin = 0;
      r_end = H;
    } else {
      const u32 u2 = unit - (u32)p.tier1_units;
      panel = u2 % p.panels;
      const int t1 = u2 / p.panels;
      plane_id = p.tier1_planes + t1 / p.strips;
      r_begin = (t1 % p.strips) * p.rows_per_strip;
      r_end = min(H, r_begin + p.rows_per_strip);
    }
    const int b = plane_id / C, c = plane_id % C;
    const bool is_anchor = c < p.M;
    const CUtensorMap* tmap = is_anchor ? &tm_anchor : &tm_part;
    const int csel = is_anchor ? c : c - p.M;
    const int K = is_anchor ? p.K : p.P;
    const int x0 = panel * kPanel - kPx;  // first column of the tile; tensor-map coordinates count 4-byte units
    const int xc = DT == SDNET_DTYPE_F32 ? x0 : x0 >> 1;
    const bool edge = S == 2 && (x0 < 0 || x0 + TileGeom<DT>::kCols > W);
    const int nrows = r_end - r_begin;
    const int groups = (nrows + 2 * R + kGroupRows - 1) / kGroupRows;  // tiles of the unit
    const int groups_out = (nrows + kGroupRows - 1) / kGroupRows;
    u64* __restrict__ list = p.lists + (size_t)plane_id * p.cap;
    int* count_ptr = p.counts + plane_id;
    int* gfloor_ptr = p.gfloor + plane_id;
    SharedFloors sf;
    sf.cta_hist = nullptr; sf.cta_floor = nullptr;
    sf.ghist = p.ghist + (size_t)plane_id * kFineBins;
    sf.gfloor = gfloor_ptr;

    UnitState st;
    st.floorx = shared_floor<DT>(__ldcg(gfloor_ptr), xscale);
#ifdef SDNET_X_NOSLOW  // timing experiment only: stream the planes, never take the slow path
    st.floorx = CUDART_INF_F;
#endif
    st.emitted = 0;
    st.nbuf = 0;
    __syncwarp();  // everyone is done with the previous unit's ring, histogram and buffer
    *reinterpret_cast<uint4*>(hist + 4 * lane) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<int4*>(minx + 4 * lane) = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff);
    __syncwarp();

    // tile k of the unit = image rows r_begin - R + 4k ..; running number tile_n + k
    auto issue = [&](u32 s, int y) {  // lane 0: pull the tile whose first image row is y into slot s
      mbar_arrive_expect_tx(bars_s + 8 * s, kTileBytes);
      if (S == 1) {
        tma_tile_4d(ring_s + s * kSlotB, tmap, xc, y, csel, b, bars_s + 8 * s);
      } else {  // y is even (r_begin even, R = 2): rows y, y+2 then rows y+1, y+3
        tma_tile_4d(ring_s + s * kSlotB, tmap, xc, y >> 1, csel, b, bars_s + 8 * s);
        tma_tile_4d(ring_s + s * kSlotB + kOddBoxOff, tmap, xc + p.odd_x - kOddShiftB / 4, y >> 1, csel, b, bars_s + 8 * s);
      }
    };
    auto wait_tile = [&](u32 n) {
      mbar_wait(bars_s + 8 * (n & (NG - 1)), (n >> 2) & 1u);
      if (edge) {  // warp-uniform
        tile_fix_edges<DT>(ring_s + (n & (NG - 1)) * kSlotB, x0, W, lane);
        __syncwarp();
      }
    };
    int y_next = r_begin - R;  // first image row of the next tile to issue
    if (lane == 0) {
      const int first = min(NG, groups);
      for (int k = 0; k < first; ++k) issue((tile_n + (u32)k) & (NG - 1), y_next + kGroupRows * k);
    }
    y_next += kGroupRows * NG;
    int gfloor_seen = 0;
    constexpr int poll_mask = 3;  // measured at 128-row strips: polling every group 0.179 ms, every 4th 0.136 ms, every 8th 0.146 ms
    u32 idx0 = (u32)(r_begin * W + panel * kPanel);  // flat index of the group's row 0, panel column 0
    wait_tile(tile_n);
    for (int g = 0; g < groups_out; ++g, idx0 += (u32)(kGroupRows * W)) {
      const u32 n = tile_n + (u32)g;  // tile holding the group's first window row
      if (R == 2 || g + 1 < groups) wait_tile(n + 1);
      if ((g & poll_mask) == 0) {
        // every 16 rows: apply the plane-wide floor fetched one period ago
        // and start the next fetch.  The load writes straight into the register it will be read
        // from a period later, so its latency is never waited for.
        st.floorx = fmaxf(st.floorx, shared_floor<DT>(gfloor_seen, xscale));
        asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(gfloor_seen) : "l"(gfloor_ptr) : "memory");
      }
      // window rows of group row i are ring rows row0 + i .. row0 + i + 2R; its centre row is row0 + i + R
      const u32 row0 = (n * kGroupRows) & kRowMask;
      uint4 c0, c1, c2, c3;
      if (R == 2) {  // centres: rows 2, 3 of this tile's slot and rows 0, 1 of the next
        const u32 a01 = ring_own + (n & (NG - 1)) * kSlotB, a23 = ring_own + ((n + 1) & (NG - 1)) * kSlotB;
        if (S == 1) {
          c0 = lds128u(a01 + 2 * kTilePitchB); c1 = lds128u(a01 + 3 * kTilePitchB);
          c2 = lds128u(a23); c3 = lds128u(a23 + kTilePitchB);
        } else {  // a slot holds rows 0, 2 in its first box and rows 1, 3 in its second
          c0 = lds128u(a01 + kTilePitchB); c1 = lds64x2(a01 + kOddBoxOff + kOddShiftB + kTilePitchB);
          c2 = lds128u(a23); c3 = lds64x2(a23 + kOddBoxOff + kOddShiftB);
        }
      } else {  // S == 1
        const u32 a012 = ring_own + (row0 + 1) * kTilePitchB, a3 = ring_own + ((row0 + 4) & kRowMask) * kTilePitchB;
        c0 = lds128u(a012); c1 = lds128u(a012 + kTilePitchB); c2 = lds128u(a012 + 2 * kTilePitchB); c3 = lds128u(a3);
      }
      if (__any_sync(0xffffffffu, TileMax<DT>::group(c0, c1, c2, c3) > st.floorx)) {
        // Something in these four rows beats the floor.  Pixel-centric slow path: (1) every (row, lane)
        // whose word of centre pixels holds one above the floor goes on the warp's work list, one
        // ballot per row, row-major; (2) settle_entries gives each listed pixel a lane of its own.
        // In the last, partial group of a strip the rows past its end belong to the next strip: they
        // may raise this alarm for nothing but are never listed.
        const float floorx = st.floorx;
        const int rows_here = nrows - g * kGroupRows;
        int nent = 0;
#pragma unroll
        for (int i = 0; i < kGroupRows; ++i) {
          const uint4 ci = i == 0 ? c0 : (i == 1 ? c1 : (i == 2 ? c2 : c3));
          const bool mine = TileMax<DT>::word(ci) > floorx && i < rows_here;
          const u32 bm = __ballot_sync(0xffffffffu, mine);
          if (mine) work[nent + __popc(bm & lt)] = (unsigned char)((i << 5) | lane);
          nent += __popc(bm);
        }
        __syncwarp();
        settle_entries<R, DT, S>(st, work, nent, ring_s, row0, floorx, idx0, W, pre, buf, hist, minx, sf, count_ptr, list,
                                 p.cap, K, lane, xscale, satx);
        // while the plane has no floor yet, publish early and often; later only in batches
        if (st.nbuf >= 16 || (st.nbuf > 0 && gfloor_seen <= 0)) {
          __syncwarp();
          flush_candidates<DT>(st, buf, hist, minx, sf, count_ptr, list, p.cap, K, lane, pre, xscale, satx);
        }
      }
      // every lane's reads of the group's first tile are done (the votes above): refill its slot
      // with the tile NG ahead
      __syncwarp();
      if (lane == 0 && g + NG < groups) {
        if (edge) fence_proxy_async();  // the slot was patched with ordinary stores
        issue(n & (NG - 1), y_next);
      }
      y_next += kGroupRows;
    }
    tile_n += (u32)groups;
    if (st.nbuf) {
      __syncwarp();
      flush_candidates<DT>(st, buf, hist, minx, sf, count_ptr, list, p.cap, K, lane, pre, xscale, satx);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// exact-select kernel: bounded-memory fallback for planes whose candidate list overflowed
// (or for every plane under SDNET_FLAG_EXACT_SELECT).  One CTA per plane: a 3-level radix
// select (11/11/10 bits) over the NMS'd scores recomputed straight from the heat map, then an
// index-ordered emission pass that keeps every score above the K-th and the lowest-index
// members of the K-th score's tie run.  Rewrites the plane's list with <= K records.
// ---------------------------------------------------------------------------------------------
constexpr int kExactThreads = 512;

struct ExactParams {
  View4 anchor, part;
  int B, M, N, H, W, K, P;
  int radius, cap, force, pre_activated;
  u64* lists;
  int* counts;
  int* flags;
};

template <int DT>
__device__ __forceinline__ u32 exact_key(const void* plane, long long sh, int H, int W, int R, int y, int x, bool pre) {
  const float v = ld_in<DT>(plane, (long long)y * sh + x);
  if (pre) {
    const u32 bits = __float_as_uint(v);
    return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
  }
  float h = v;
  for (int dy = -R; dy <= R; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -R; dx <= R; ++dx) {
      const int xx = x + dx;
      if (xx >= 0 && xx < W) h = fmaxf(h, ld_in<DT>(plane, (long long)yy * sh + xx));
    }
  }
  const float sv = Num<DT>::act(v);
  const bool peak = (v == h) || (sv == Num<DT>::act(h));
  return peak ? __float_as_uint(sv) : 0u;
}

// key of a pixel already known to survive NMS
template <int DT>
__device__ __forceinline__ u32 survivor_key(const void* plane, long long idx, bool pre) {
  const float v = ld_in<DT>(plane, idx);
  if (pre) {
    const u32 bits = __float_as_uint(v);
    return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
  }
  return __float_as_uint(Num<DT>::act(v));
}

// digit (from the top) at which the cumulative count reaches `need`; bins = 2048
__device__ void pick_digit(const u32* s_hist, int nbins, int need, int* s_out) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int per = nbins / 32;
    u32 s = 0;
    for (int q = 0; q < per; ++q) s += s_hist[lane * per + q];
    u32 suf = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      u32 t = __shfl_down_sync(0xffffffffu, suf, d);
      if (lane + d < 32) suf += t;
    }
    const u32 mask = __ballot_sync(0xffffffffu, suf >= (u32)need);
    if (mask == 0) {
      if (lane == 0) { s_out[0] = -1; s_out[1] = 0; }
    } else {
      const int L = 31 - __clz(mask);
      if (lane == L) {
        u32 above = suf - s;
        int dsel = L * per;
        for (int q = per - 1; q >= 0; --q) {
          const u32 cq = s_hist[L * per + q];
          if (above + cq >= (u32)need) { dsel = L * per + q; break; }
          above += cq;
        }
        s_out[0] = dsel;
        s_out[1] = (int)above;
      }
    }
  }
  __syncthreads();
}

template <int DT>
__global__ void __launch_bounds__(kExactThreads) sdnet_exact_select_kernel(const __grid_constant__ ExactParams p) {
  __shared__ u32 s_hist[2048];
  __shared__ int s_out[2];
  __shared__ int s_warp[kExactThreads / 32][2];
  __shared__ int s_base[2];
  const int C = p.M + p.N;
  const int planes = p.B * C;
  pdl_launch_dependents();
  pdl_wait();  // the peaks kernel's lists and counts are complete and visible
  if (!p.force) {
    // common case: nothing overflowed.  One coalesced look at this CTA's planes, then leave.
    int mine = 0;
    for (int q = blockIdx.x + threadIdx.x * gridDim.x; q < planes; q += blockDim.x * gridDim.x)
      mine |= p.counts[q] > p.cap;
    if (!__syncthreads_or(mine)) return;
  }
  for (int plane_id = blockIdx.x; plane_id < planes; plane_id += gridDim.x) {
  const int emitted = p.counts[plane_id];
  if (!p.force && emitted <= p.cap) continue;  // block-uniform
  __syncthreads();
  const int b = plane_id / C, c = plane_id % C;
  const bool is_anchor = c < p.M;
  const View4& vw = is_anchor ? p.anchor : p.part;
  const void* plane = static_cast<const typename Num<DT>::In*>(vw.data) + (long long)b * vw.sb +
                      (long long)(is_anchor ? c : c - p.M) * vw.sc;
  const long long sh = vw.sh;
  const int K = is_anchor ? p.K : p.P;
  const int H = p.H, W = p.W, HW = H * W, R = p.radius;
  const bool pre = p.pre_activated != 0;
  const int tid = threadIdx.x;
  u64* list = p.lists + (size_t)plane_id * p.cap;
  // per-pixel key cache lives behind the K output records of this plane's list region
  // (cap*8 bytes >= K*8 + H*W/8*... see plan_workspace): 1 bit per pixel "survives NMS".
  u32* bitmap = reinterpret_cast<u32*>(list + K + 2);
  const int words = (HW + 31) / 32;

  // level 1 (top 11 bits) + NMS bitmap
  for (int i = tid; i < 2048; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  for (int base = 0; base < words * 32; base += blockDim.x) {
    const int i = base + tid;
    u32 key = 0;
    if (i < HW) key = exact_key<DT>(plane, sh, H, W, R, i / W, i % W, pre);
    const u32 m = __ballot_sync(0xffffffffu, key != 0);
    if ((tid & 31) == 0 && (i >> 5) < words) bitmap[i >> 5] = m;
    if (key) atomicAdd(&s_hist[key >> 21], 1u);
  }
  __syncthreads();
  u32 prefix = 0;
  int need = K;
  bool all = false;
  pick_digit(s_hist, 2048, need, s_out);
  if (s_out[0] < 0) all = true;  // fewer than K survivors: keep them all
  u32 thresh = 0;
  if (!all) {
    need -= s_out[1];
    prefix = (u32)s_out[0];
    __syncthreads();
    // level 2 (next 11 bits)
    for (int i = tid; i < 2048; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < HW; i += blockDim.x) {
      if (!((bitmap[i >> 5] >> (i & 31)) & 1u)) continue;
      const u32 key = survivor_key<DT>(plane, (long long)(i / W) * sh + (i % W), pre);
      if ((key >> 21) == prefix) atomicAdd(&s_hist[(key >> 10) & 0x7ffu], 1u);
    }
    __syncthreads();
    pick_digit(s_hist, 2048, need, s_out);
    need -= s_out[1];
    prefix = (prefix << 11) | (u32)s_out[0];
    __syncthreads();
    // level 3 (last 10 bits)
    for (int i = tid; i < 1024; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < HW; i += blockDim.x) {
      if (!((bitmap[i >> 5] >> (i & 31)) & 1u)) continue;
      const u32 key = survivor_key<DT>(plane, (long long)(i / W) * sh + (i % W), pre);
      if ((key >> 10) == prefix) atomicAdd(&s_hist[key & 0x3ffu], 1u);
    }
    __syncthreads();
    pick_digit(s_hist, 1024, need, s_out);
    need -= s_out[1];  // members of the K-th score's tie run still to take, lowest index first
    thresh = (prefix << 10) | (u32)s_out[0];
    __syncthreads();
  }
  // ordered emission
  if (tid < 2) s_base[tid] = 0;  // [0] scores above the threshold so far, [1] tie-run members so far
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int base = 0; base < HW; base += blockDim.x) {
    const int i = base + tid;
    u32 key = 0;
    if (i < HW && ((bitmap[i >> 5] >> (i & 31)) & 1u)) {
      key = survivor_key<DT>(plane, (long long)(i / W) * sh + (i % W), pre);
    }
    const bool gt = key != 0 && (all || key > thresh);
    const bool eq = key != 0 && !all && key == thresh;
    const u32 mg = __ballot_sync(0xffffffffu, gt), me = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { s_warp[warp][0] = __popc(mg); s_warp[warp][1] = __popc(me); }
    __syncthreads();
    int g_before = 0, e_before = 0;
    for (int w = 0; w < warp; ++w) { g_before += s_warp[w][0]; e_before += s_warp[w][1]; }
    const u32 lt = (1u << lane) - 1u;
    const int e_rank = s_base[1] + e_before + __popc(me & lt);  // tie-run members with a lower index
    const bool take_eq = eq && e_rank < need;
    const int pos = s_base[0] + g_before + __popc(mg & lt) + min(e_rank, need);
    if (gt || take_eq) list[pos] = ((u64)key << 32) | (u32)i;
    __syncthreads();
    if (tid == 0) {
      int g = 0, e = 0;
      for (int w = 0; w < kExactThreads / 32; ++w) { g += s_warp[w][0]; e += s_warp[w][1]; }
      s_base[0] += g;  // scores above the threshold so far
      s_base[1] += e;  // tie-run members so far
    }
    __syncthreads();
  }
  if (tid == 0) {
    p.counts[plane_id] = s_base[0] + min(s_base[1], need);
    p.flags[plane_id] = 1;
  }
  __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// tail kernel: select + sort + gather + group, one CTA per image
// ---------------------------------------------------------------------------------------------
struct TailParams {
  View4 offsets, embeddings;
  int B, M, N, H, W, K, P;
  int cap;
  int pre_activated, no_grouping;
  float conf, dist_abs;
  const u64* lists;
  const int* counts;
  float* anchor_out;
  float* part_out;
  long long* anchor_inds;
  long long* part_inds;
  float* part_emb;
  int* assign;
  int* out_counts;
  int* diag;
  const int* exact_flags;  // [planes] 1 if the exact select rewrote the list
  int n_dest;              // fused gather: every output is stored n_dest times, at ptr + dest_delta[j]
  long long dest_delta[SDNET_MAX_DEST];
};

// Store one output value locally (n_dest == 0) or into every destination copy of the output blob
// (fused detection gather: peer-mapped symmetric memory, plain st.global over NVLink).
template <typename T>
__device__ __forceinline__ void store_out(const TailParams& p, T* ptr, const T& v) {
  if (p.n_dest == 0) {
    *ptr = v;
  } else {
    for (int j = 0; j < p.n_dest; ++j) *reinterpret_cast<T*>(reinterpret_cast<char*>(ptr) + p.dest_delta[j]) = v;
  }
}

__device__ __forceinline__ u64 make_comp(u64 rec, int c_local) {
  // rec = key:32 | idx:32  ->  key:32 | (255-c):8 | (0xFFFFFF-idx):24 ; larger = earlier in the output
  const u32 key = (u32)(rec >> 32);
  const u32 idx = (u32)rec;
  return ((u64)key << 32) | ((u64)(255u - (u32)c_local) << 24) | (u64)(0xFFFFFFu - idx);
}

// The tail CTA is split into two teams of kTeamThreads threads that work concurrently: team 0
// selects the anchors, team 1 the parts; each has its own sort buffer and syncs on its own named
// barrier.  They meet once, before the grouping.
constexpr int kTeamThreads = 256;

struct Team {
  int tid;    // thread index inside the team
  int id;     // 0 = anchors, 1 = parts
  __device__ __forceinline__ void sync() const {
    asm volatile("bar.sync %0, %1;" ::"r"(id + 1), "r"(kTeamThreads) : "memory");
  }
};

__device__ void bitonic_sort_desc(const Team& tm, u64* s, int n /* power of two */) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tm.tid; i < n; i += kTeamThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const u64 a = s[i], b = s[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) { s[i] = b; s[ixj] = a; }
        }
      }
      tm.sync();
    }
  }
}

// Select the `want` largest composites of planes [c0, c0+nc) of image b into s_sel (sorted
// descending).  Returns the number selected (< want only when fewer candidates exist).
__device__ int select_group(const Team& tm, const TailParams& p, int b, int c0, int nc, int want, u64* s_sel,
                            u32* s_hist, int* s_misc) {
  const int C = p.M + p.N;
  const int tid = tm.tid;
  // total candidates
  if (tid == 0) {
    int tot = 0;
    for (int c = 0; c < nc; ++c) tot += min(p.counts[(size_t)b * C + c0 + c], p.cap);
    s_misc[0] = tot;
    s_misc[1] = 0;  // collected
  }
  tm.sync();
  const int total = s_misc[0];
  u64 prefix = 0;   // value of the top `bits` bits that boundary elements share
  int bits = 0;
  // radix-refine until what is left (everything certainly selected + the boundary bucket) is a
  // small sort: the bitonic network below costs O(n log^2 n) and dominated this kernel when it
  // was handed the full 2048-element buffer
  const int target = min(kSortN, max(want + 64, 128));
  if (total > target) {
    int need = want;      // how many still have to come from the boundary bucket
    int certain = 0;      // elements strictly above the boundary bucket
    for (int level = 0; level < 8; ++level) {
      const int shift = 56 - 8 * level;
      for (int i = tid; i < 256; i += kTeamThreads) s_hist[i] = 0;
      tm.sync();
      for (int c = 0; c < nc; ++c) {
        const int n = min(p.counts[(size_t)b * C + c0 + c], p.cap);
        const u64* list = p.lists + ((size_t)b * C + c0 + c) * p.cap;
        for (int i = tid; i < n; i += kTeamThreads) {
          const u64 v = make_comp(list[i], c);
          if (bits == 0 || (v >> (64 - bits)) == prefix) atomicAdd(&s_hist[(u32)(v >> shift) & 0xffu], 1u);
        }
      }
      tm.sync();
      if (tid < 32) {
        // lane l owns digits 8l..8l+7; suffix-scan from the top
        u32 loc[8], sum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { loc[q] = s_hist[8 * tid + q]; sum += loc[q]; }
        u32 suf = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          u32 t = __shfl_down_sync(0xffffffffu, suf, d);
          if (tid + d < 32) suf += t;
        }
        const u32 mask = __ballot_sync(0xffffffffu, suf >= (u32)need);
        const int L = 31 - __clz(mask);  // mask != 0 because the bucket holds >= need elements
        if (tid == L) {
          u32 above = suf - sum;
          int dsel = 8 * L;
          for (int q = 7; q >= 0; --q) {
            if (above + loc[q] >= (u32)need) { dsel = 8 * L + q; break; }
            above += loc[q];
          }
          s_misc[2] = dsel;
          s_misc[3] = (int)above;                 // elements in this bucket with a larger digit
          s_misc[4] = (int)s_hist[dsel];          // size of the new boundary bucket
        }
      }
      tm.sync();
      const int dsel = s_misc[2], above = s_misc[3], binc = s_misc[4];
      certain += above;
      need -= above;
      prefix = (prefix << 8) | (u64)dsel;
      bits += 8;
      tm.sync();
      if (certain + binc <= target) break;  // everything at or above the boundary bucket is a small sort
    }
  }
  // collect: all elements whose top `bits` bits are >= prefix
  for (int c = 0; c < nc; ++c) {
    const int n = min(p.counts[(size_t)b * C + c0 + c], p.cap);
    const u64* list = p.lists + ((size_t)b * C + c0 + c) * p.cap;
    for (int i = tid; i < n; i += kTeamThreads) {
      const u64 v = make_comp(list[i], c);
      if (bits == 0 || (v >> (64 - bits)) >= prefix) {
        const int slot = atomicAdd(&s_misc[1], 1);
        if (slot < kSortN) s_sel[slot] = v;
      }
    }
  }
  tm.sync();
  const int got = min(s_misc[1], kSortN);
  int n2 = 32;
  while (n2 < got) n2 <<= 1;
  for (int i = got + tid; i < n2; i += kTeamThreads) s_sel[i] = 0;
  tm.sync();
  bitonic_sort_desc(tm, s_sel, n2);
  return min(got, want);
}

// Slots [have, want) of a group are the zero-valued entries torch.topk pads with: under
// (value desc, index asc) they are the lowest-index pixels of the group's first plane that
// did not survive NMS.  All of them lie below index `want`.
__device__ void zero_fill(const Team& tm, const TailParams& p, int b, int c0, int have, int want, u64* s_sel,
                          u32* s_bits) {
  const int C = p.M + p.N;
  const int tid = tm.tid;
  const int words = (want + 31) / 32;
  for (int i = tid; i < words; i += kTeamThreads) s_bits[i] = 0;
  tm.sync();
  const int n = min(p.counts[(size_t)b * C + c0], p.cap);
  const u64* list = p.lists + ((size_t)b * C + c0) * p.cap;
  for (int i = tid; i < n; i += kTeamThreads) {
    const u32 idx = (u32)list[i];
    if (idx < (u32)want) atomicOr(&s_bits[idx >> 5], 1u << (idx & 31));
  }
  tm.sync();
  for (int s = have + tid; s < want; s += kTeamThreads) {
    int rank = s - have;  // rank-th non-peak index
    int w = 0;
    for (; w < words; ++w) {
      const int z = 32 - __popc(s_bits[w]);
      if (rank < z) break;
      rank -= z;
    }
    const u32 free_mask = ~s_bits[w];
    const u32 bit = __fns(free_mask, 0, rank + 1);
    const u32 idx = (u32)(w * 32) + bit;
    // key 0 (score 0.0), class 0
    s_sel[s] = ((u64)255u << 24) | (u64)(0xFFFFFFu - idx);
  }
  tm.sync();
}

__device__ __forceinline__ float key_to_score(u32 key, bool pre) {
  if (!pre) return __uint_as_float(key);
  const u32 bits = (key == 0) ? 0u : ((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
  return __uint_as_float(bits);
}

template <int DT>
__global__ void __launch_bounds__(2 * kTeamThreads) sdnet_tail_kernel(const __grid_constant__ TailParams p) {
  __shared__ u64 s_sel[2][kSortN];
  __shared__ u32 s_hist[2][256];
  __shared__ int s_misc[2][8];
  __shared__ float s_ax[SDNET_MAX_TOPK], s_ay[SDNET_MAX_TOPK];
  __shared__ int s_cnt[2];
  const int b = blockIdx.x;
  Team tm;
  tm.id = threadIdx.x / kTeamThreads;
  tm.tid = threadIdx.x % kTeamThreads;
  const int W = p.W;
  const bool pre = p.pre_activated != 0;
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  pdl_wait();  // candidate lists (peaks kernel, possibly rewritten by the exact select) are final
  __syncthreads();
  typedef typename Num<DT>::In In;
  const In* offx = static_cast<const In*>(p.offsets.data) + (long long)b * p.offsets.sb;
  const In* offy = offx + p.offsets.sc;
  const long long osh = p.offsets.sh;
  u64* sel = s_sel[tm.id];

  if (tm.id == 0) {
    // ---- anchors
    const int have = select_group(tm, p, b, 0, p.M, p.K, sel, s_hist[0], s_misc[0]);
    if (have < p.K) zero_fill(tm, p, b, 0, have, p.K, sel, s_hist[0]);
    int n_valid = 0;
    for (int s = tm.tid; s < p.K; s += kTeamThreads) {
      const u64 v = sel[s];
      const int cls = 255 - (int)((v >> 24) & 0xffu);
      const u32 idx = 0xFFFFFFu - (u32)(v & 0xFFFFFFu);
      const int yy = idx / W, xx = idx - yy * W;
      const float score = key_to_score((u32)(v >> 32), pre);
      const float x = __fadd_rn((float)xx, Num<DT>::to_float(__ldg(offx + (long long)yy * osh + xx)));
      const float y = __fadd_rn((float)yy, Num<DT>::to_float(__ldg(offy + (long long)yy * osh + xx)));
      store_out(p, reinterpret_cast<float4*>(p.anchor_out) + (size_t)b * p.K + s, make_float4(x, y, score, (float)cls));
      store_out(p, p.anchor_inds + (size_t)b * p.K + s, (long long)idx);
      const bool valid = score > p.conf;
      // masked anchors sit at (+1e6, +1e6): decoders.py:85-86
      s_ax[s] = valid ? x : kFar;
      s_ay[s] = valid ? y : kFar;
      n_valid += valid ? 1 : 0;
    }
    if (n_valid) atomicAdd(&s_cnt[0], n_valid);
  } else {
    // ---- parts
    const int have = select_group(tm, p, b, p.M, p.N, p.P, sel, s_hist[1], s_misc[1]);
    if (have < p.P) zero_fill(tm, p, b, p.M, have, p.P, sel, s_hist[1]);
    const In* embx = p.embeddings.data ? static_cast<const In*>(p.embeddings.data) + (long long)b * p.embeddings.sb : nullptr;
    const In* emby = embx ? embx + p.embeddings.sc : nullptr;
    const long long esh = p.embeddings.sh;
    int n_valid = 0;
    for (int s = tm.tid; s < p.P; s += kTeamThreads) {
      const u64 v = sel[s];
      const int cls = 255 - (int)((v >> 24) & 0xffu);
      const u32 idx = 0xFFFFFFu - (u32)(v & 0xFFFFFFu);
      const int yy = idx / W, xx = idx - yy * W;
      const float score = key_to_score((u32)(v >> 32), pre);
      const float x = __fadd_rn((float)xx, Num<DT>::to_float(__ldg(offx + (long long)yy * osh + xx)));
      const float y = __fadd_rn((float)yy, Num<DT>::to_float(__ldg(offy + (long long)yy * osh + xx)));
      float ex = 0.f, ey = 0.f;
      if (embx) {
        ex = Num<DT>::to_float(__ldg(embx + (long long)yy * esh + xx));
        ey = Num<DT>::to_float(__ldg(emby + (long long)yy * esh + xx));
      }
      const float ox = __fadd_rn(x, ex), oy = __fadd_rn(y, ey);
      float2* po = reinterpret_cast<float2*>(p.part_out + ((size_t)b * p.P + s) * 6);
      store_out(p, po + 0, make_float2(x, y));
      store_out(p, po + 1, make_float2(score, (float)cls));
      store_out(p, po + 2, make_float2(ox, oy));
      store_out(p, p.part_inds + (size_t)b * p.P + s, (long long)idx);
      if (p.part_emb) store_out(p, reinterpret_cast<float2*>(p.part_emb) + (size_t)b * p.P + s, make_float2(ex, ey));
      const bool valid = score > p.conf;
      n_valid += valid ? 1 : 0;
      // masked parts sit at (-1e6, -1e6): decoders.py:80-81.  The slot's composite is no longer
      // needed: keep the part's origin there for the grouping pass.
      reinterpret_cast<float2*>(sel)[s] = make_float2(valid ? ox : -kFar, valid ? oy : -kFar);
    }
    if (n_valid) atomicAdd(&s_cnt[1], n_valid);
  }
  __syncthreads();

  // ---- grouping: every part to its nearest anchor (first minimum), gated by the distance threshold
  // (reference: hypot utils.py:422-437, min(dim=1) decoders.py:99).  A part is shared by g lanes of one
  // warp, lane `sub` taking anchors sub, sub + g, ...  The square root is monotone, so the smallest
  // distance is the root of the smallest squared distance m2 -- but two different squares can round to
  // the same root and the reference's min() then keeps the FIRST anchor.  Hence two sweeps without a
  // root in the loop: m2, then the first anchor whose square lies within 1e-6 of m2 (a root can only
  // tie if its square is within 2^-22 relative) AND whose root equals root(m2).
  const float2* origin = reinterpret_cast<const float2*>(s_sel[1]);
  int g = 32;
  while (g > 1 && p.P * g > (int)blockDim.x) g >>= 1;
  const int sub = threadIdx.x & (g - 1), per_pass = blockDim.x / g;
  for (int s0 = 0; s0 < p.P; s0 += per_pass) {  // block-uniform
    const int s = s0 + (int)threadIdx.x / g;
    const bool live = s < p.P;
    int slot = -1;
    if (!p.no_grouping) {  // kernel-uniform
      const float qx = live ? origin[s].x : 0.f, qy = live ? origin[s].y : 0.f;
      float m2 = CUDART_INF_F;
      for (int a = sub; a < p.K; a += g) {
        const float dx = __fsub_rn(qx, s_ax[a]), dy = __fsub_rn(qy, s_ay[a]);
        m2 = fminf(m2, __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      }
      for (int w = g >> 1; w > 0; w >>= 1) m2 = fminf(m2, __shfl_xor_sync(0xffffffffu, m2, w));
      const float best = __fsqrt_rn(m2), near = m2 * 1.000001f;
      int arg = 0x7fffffff;
      for (int a = sub; a < p.K; a += g) {
        const float dx = __fsub_rn(qx, s_ax[a]), dy = __fsub_rn(qy, s_ay[a]);
        const float sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (sq <= near && __fsqrt_rn(sq) == best) { arg = a; break; }
      }
      for (int w = g >> 1; w > 0; w >>= 1) arg = min(arg, __shfl_xor_sync(0xffffffffu, arg, w));
      slot = (best < p.dist_abs) ? arg : -1;
    }
    if (live && sub == 0) store_out(p, p.assign + (size_t)b * p.P + s, slot);
  }
  if (threadIdx.x < 2) store_out(p, p.out_counts + (size_t)b * 2 + threadIdx.x, s_cnt[threadIdx.x]);
  if (p.diag) {
    const int C = p.M + p.N;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      store_out(p, p.diag + ((size_t)b * C + c) * 2 + 0, p.counts[(size_t)b * C + c]);
      store_out(p, p.diag + ((size_t)b * C + c) * 2 + 1, p.exact_flags[(size_t)b * C + c]);
    }
  }
  if (p.n_dest) __threadfence_system();  // peer stores performed before the kernel retires
}

// ---------------------------------------------------------------------------------------------
// metadata: clamped-sigmoid maps
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void sdnet_activate_kernel(View4 in, int C, int H, int W, size_t total, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    size_t t = i / W;
    const int y = (int)(t % H);
    t /= H;
    const int c = (int)(t % C);
    const long long b = (long long)(t / C);
    out[i] = Num<DT>::act(ld_in<DT>(in.data, b * in.sb + (long long)c * in.sc + (long long)y * in.sh + x));
  }
}

// ---------------------------------------------------------------------------------------------
// evaluator matching (the step after the path): reference Evaluator.eval_anchor / eval_part,
// src/sdnet/model/evaluator.py:244-334, on the packed detections.  One CTA per image; the greedy
// "first detection in score order that claims a ground truth wins it" loop is order-free once every
// detection knows its nearest ground truth: the winner of ground truth j is the smallest slot index
// among the detections whose nearest is j and whose distance is under the threshold (atomicMin).
// Arithmetic is double throughout, like the reference's Python floats.
// ---------------------------------------------------------------------------------------------
constexpr int kMatchThreads = 256;

__device__ void match_pass(const float* __restrict__ out, int row_len, int slots, int n_classes, bool strict_gt,
                           double conf, double sx, double sy, const double* __restrict__ scale,
                           const double* __restrict__ gt, int n_gt, int* __restrict__ stats_out,
                           double* __restrict__ acc_out, double* s_gx, double* s_gy, int* s_glab, int* s_winner,
                           int* s_jmin, double* s_dist, int* s_stats) {
  const double rx = scale[0], ry = scale[1], thresh = scale[2], norm = scale[3];
  for (int j = threadIdx.x; j < n_gt; j += blockDim.x) {
    s_gx[j] = gt[3 * j + 0] * rx;  // annotation.resized(...): evaluator.py:247
    s_gy[j] = gt[3 * j + 1] * ry;
    s_glab[j] = (int)gt[3 * j + 2];
    s_winner[j] = 0x7fffffff;
  }
  for (int i = threadIdx.x; i < 3 * n_classes; i += blockDim.x) s_stats[i] = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < n_gt; j += blockDim.x)
    if (s_glab[j] >= 0 && s_glab[j] < n_classes) atomicAdd(&s_stats[3 * s_glab[j] + 1], 1);  // npos
  for (int i = threadIdx.x; i < slots; i += blockDim.x) {
    const float* row = out + (size_t)i * row_len;
    const double score = (double)row[2];
    const bool det = strict_gt ? (score > conf) : !(score < conf);
    const int lab = (int)row[3];
    int jmin = -1;
    double best = 1.7976931348623157e308;  // sys.float_info.max
    if (det) {
      atomicAdd(&s_stats[3 * lab + 0], 1);  // ndet
      const double px = ((double)row[0] * sx) * rx, py = ((double)row[1] * sy) * ry;  // decoders.py:139 then evaluator.py:248
      for (int j = 0; j < n_gt; ++j) {
        if (s_glab[j] != lab) continue;
        const double d = hypot(px - s_gx[j], py - s_gy[j]);  // np.hypot, utils.py:31-32
        if (d < best) { best = d; jmin = j; }
      }
      if (jmin >= 0 && best < thresh) atomicMin(&s_winner[jmin], i);
    }
    s_jmin[i] = det ? jmin : -1;
    s_dist[i] = best;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < slots; i += blockDim.x) {
    const int jmin = s_jmin[i];
    const bool tp = jmin >= 0 && s_dist[i] < thresh && s_winner[jmin] == i;
    acc_out[i] = tp ? s_dist[i] / norm : __longlong_as_double(0x7ff8000000000000ll);
    if (tp) atomicAdd(&s_stats[3 * (int)out[(size_t)i * row_len + 3] + 2], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * n_classes; i += blockDim.x) stats_out[i] = s_stats[i];
  __syncthreads();
}

__global__ void __launch_bounds__(kMatchThreads) sdnet_match_kernel(const SdnetMatchParams p) {
  __shared__ double s_gx[SDNET_MAX_GT], s_gy[SDNET_MAX_GT], s_dist[SDNET_MAX_TOPK];
  __shared__ int s_glab[SDNET_MAX_GT], s_winner[SDNET_MAX_GT], s_jmin[SDNET_MAX_TOPK], s_stats[3 * SDNET_MAX_CHANNELS];
  const int b = blockIdx.x;
  const double* scale = p.image_scale + 4 * (size_t)b;
  match_pass(p.anchor_out + (size_t)b * p.K * 4, 4, p.K, p.M, true, p.conf, p.sx, p.sy, scale,
             p.gt_anchors + (size_t)b * p.max_gt_anchors * 3, min(p.n_gt_anchors[b], p.max_gt_anchors),
             p.anchor_stats + (size_t)b * p.M * 3, p.anchor_acc + (size_t)b * p.K, s_gx, s_gy, s_glab, s_winner, s_jmin,
             s_dist, s_stats);
  match_pass(p.part_out + (size_t)b * p.P * 6, 6, p.P, p.N, false, p.conf, p.sx, p.sy, scale,
             p.gt_parts + (size_t)b * p.max_gt_parts * 3, min(p.n_gt_parts[b], p.max_gt_parts),
             p.part_stats + (size_t)b * p.N * 3, p.part_acc + (size_t)b * p.P, s_gx, s_gy, s_glab, s_winner, s_jmin, s_dist,
             s_stats);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Workspace {
  size_t off_counts, off_flags, off_sched, off_gfloor, off_ghist, off_lists, total;
  int cap;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Workspace plan_workspace(int B, int M, int N, int H, int W, int K, int P) {
  Workspace ws;
  const size_t planes = (size_t)B * (M + N);
  const int kmax = K > P ? K : P;
  ws.cap = (int)((size_t)H * W / 8 + 2 * (size_t)kmax + 1024);
  size_t off = 0;
  ws.off_counts = off; off = align_up(off + planes * sizeof(int), 256);
  ws.off_flags = off;  off = align_up(off + planes * sizeof(int), 256);
  ws.off_sched = off;  off = align_up(off + 64, 256);
  ws.off_gfloor = off; off = align_up(off + planes * sizeof(int), 256);
  ws.off_ghist = off;  off = align_up(off + planes * kFineBins * sizeof(u32), 256);
  ws.off_lists = off;  off = align_up(off + planes * (size_t)ws.cap * sizeof(u64), 256);
  ws.total = off;
  return ws;
}

int device_sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
  cached = sms;
  return sms;
}

int validate(const SdnetDecodeParams* p) {
  if (!p) return SDNET_E_NULL;
  if (p->struct_size != sizeof(SdnetDecodeParams)) return SDNET_E_STRUCT;
  if (p->dtype != SDNET_DTYPE_F32 && p->dtype != SDNET_DTYPE_F16 && p->dtype != SDNET_DTYPE_BF16) return SDNET_E_DTYPE;
  if (p->B <= 0 || p->M <= 0 || p->N <= 0 || p->H <= 0 || p->W <= 0 || p->K <= 0 || p->P <= 0) return SDNET_E_SHAPE;
  const long long hw = (long long)p->H * p->W;
  if (hw >= (1ll << 24) || p->K > hw || p->P > hw) return SDNET_E_SHAPE;
  if (p->K > SDNET_MAX_TOPK || p->P > SDNET_MAX_TOPK || p->M + p->N > SDNET_MAX_CHANNELS) return SDNET_E_SHAPE;
  if (p->radius != 1 && p->radius != 2) return SDNET_E_RADIUS;
  if (p->n_dest < 0 || p->n_dest > SDNET_MAX_DEST) return SDNET_E_SHAPE;
  const bool no_group = (p->flags & SDNET_FLAG_NO_GROUPING) != 0;
  if (!p->anchor_hm.data || !p->part_hm.data || !p->offsets.data || (!no_group && !p->embeddings.data)) return SDNET_E_NULL;
  if (!p->anchor_out || !p->part_out || !p->anchor_inds || !p->part_inds || !p->assign || !p->counts) return SDNET_E_NULL;
  if (p->anchor_hm.stride_w != 1 || p->part_hm.stride_w != 1 || p->offsets.stride_w != 1 ||
      (p->embeddings.data && p->embeddings.stride_w != 1))
    return SDNET_E_STRIDE;
  const Workspace ws = plan_workspace(p->B, p->M, p->N, p->H, p->W, p->K, p->P);
  if (!p->workspace || ((uintptr_t)p->workspace & 255) || p->workspace_bytes < ws.total) return SDNET_E_WORKSPACE;
  return 0;
}

View4 to_view(const SdnetTensor4& t) {
  View4 v;
  v.data = t.data;
  v.sb = t.stride_b; v.sc = t.stride_c; v.sh = t.stride_h;
  return v;
}

bool view_aligned(const SdnetTensor4& t, int W) {
  return ((uintptr_t)t.data % 16 == 0) && (t.stride_b % 4 == 0) && (t.stride_c % 4 == 0) && (t.stride_h % 4 == 0) &&
         (W % 4 == 0);
}

constexpr int kPeaksCtasPerSm = 4;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

// How the tile kernel can read a view: 0 = not at all (take the per-lane kernel), 1 = plain rows (pitch
// a multiple of 16 bytes), 2 = row pairs (fp16/bf16 with an 8-byte-multiple pitch, e.g. W = 612).
int tile_rows_per_tma_row(const SdnetTensor4& t, int dtype, int H, int W, int radius) {
  const int px = dtype == SDNET_DTYPE_F32 ? 4 : 8;  // elements per 16 bytes
  if ((uintptr_t)t.data % 16 != 0 || t.stride_b % px != 0 || t.stride_c % px != 0 || t.stride_h < W || W % 4 != 0) return 0;
  if (t.stride_h % px == 0) return 1;
  if (dtype != SDNET_DTYPE_F32 && t.stride_h % 8 == 4 && H % 2 == 0 && radius == 2) return 2;
  return 0;
}

// (W, H, channels, B) view -> tensor map with a (128|256 + halo) x 4 x 1 x 1 box and NaN out-of-bounds
// fill; rows_per_tma_row = 2: (pitch + W, H/2, channels, B) with a 2-row box (see the tile kernel)
bool make_tile_map(CUtensorMap* map, const SdnetTensor4& t, int dtype, int B, int Cn, int H, int W, int rows_per_tma_row) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  // A box side is limited to 256 elements and a fp16/bf16 tile row has 272, so those maps describe
  // PAIRS of elements as one 32-bit float.  Every stride and W are even there, coordinates are halved
  // in the kernel, and the out-of-bounds fill of a float32 map -- measured 0x7FF77FF7 on B200
  // (tools/probes/tma_fill.cu) -- reads as two NaNs in fp16 and in bf16 alike.
  const cuuint64_t esz = dtype == SDNET_DTYPE_F32 ? 4 : 2, per = 4 / esz;
  const int S = rows_per_tma_row;
  const cuuint64_t dims[4] = {(cuuint64_t)(S == 1 ? W : t.stride_h + W) / per, (cuuint64_t)(H / S), (cuuint64_t)Cn, (cuuint64_t)B};
  // strides of size-1 dimensions are arbitrary in torch: make them canonical
  const cuuint64_t sh = (cuuint64_t)t.stride_h * esz;
  const cuuint64_t sc = Cn > 1 ? (cuuint64_t)t.stride_c * esz : sh * (cuuint64_t)H;
  const cuuint64_t sb = B > 1 ? (cuuint64_t)t.stride_b * esz : sc * (cuuint64_t)Cn;
  const cuuint64_t strides[3] = {sh * S, sc, sb};
  const cuuint32_t box[4] = {(cuuint32_t)(kTilePitchB / 4), (cuuint32_t)(kGroupRows / S), 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(t.data), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA) == CUDA_SUCCESS;
}

// the tile kernel instantiation for (radius, dtype, rows per TMA row)
typedef void (*TileKernel)(const PeaksParams, const CUtensorMap, const CUtensorMap);
TileKernel tile_kernel_for(int radius, int dtype, int S) {
  if (dtype == SDNET_DTYPE_F32) return radius == 2 ? sdnet_peaks_tile_kernel<2, SDNET_DTYPE_F32, 1> : sdnet_peaks_tile_kernel<1, SDNET_DTYPE_F32, 1>;
  if (dtype == SDNET_DTYPE_F16) {
    if (S == 2) return sdnet_peaks_tile_kernel<2, SDNET_DTYPE_F16, 2>;
    return radius == 2 ? sdnet_peaks_tile_kernel<2, SDNET_DTYPE_F16, 1> : sdnet_peaks_tile_kernel<1, SDNET_DTYPE_F16, 1>;
  }
  if (S == 2) return sdnet_peaks_tile_kernel<2, SDNET_DTYPE_BF16, 2>;
  return radius == 2 ? sdnet_peaks_tile_kernel<2, SDNET_DTYPE_BF16, 1> : sdnet_peaks_tile_kernel<1, SDNET_DTYPE_BF16, 1>;
}

// Which peaks kernel a decode of these tensors runs (SDNET_PATH_*), encoding the tensor maps on the way.
int select_peaks_path(const SdnetDecodeParams* p, CUtensorMap* tm_anchor, CUtensorMap* tm_part, int* tile_s_out) {
  static const int path_override = [] {  // tuning knob, read once: SDNET_PEAKS_PATH = tile | cta | warp
    const char* e = getenv("SDNET_PEAKS_PATH");
    if (!e) return 0;
    return e[0] == 't' ? 1 : (e[0] == 'c' ? 2 : (e[0] == 'w' ? 3 : 0));
  }();
  if ((p->flags & SDNET_FLAG_WARP_KERNEL) || path_override == 3) return SDNET_PATH_WARP;
  const bool is_f32 = p->dtype == SDNET_DTYPE_F32;  // the warp-specialised kernel is fp32-only
  int tile_s = tile_rows_per_tma_row(p->anchor_hm, p->dtype, p->H, p->W, p->radius);
  if (tile_s != tile_rows_per_tma_row(p->part_hm, p->dtype, p->H, p->W, p->radius) ||
      (tile_s == 2 && p->anchor_hm.stride_h != p->part_hm.stride_h))
    tile_s = 0;
  if (tile_s != 0 && path_override != 2 &&
      make_tile_map(tm_anchor, p->anchor_hm, p->dtype, p->B, p->M, p->H, p->W, tile_s) &&
      make_tile_map(tm_part, p->part_hm, p->dtype, p->B, p->N, p->H, p->W, tile_s)) {
    *tile_s_out = tile_s;
    return tile_s == 2 ? SDNET_PATH_TILE_ROW_PAIRS : SDNET_PATH_TILE;
  }
  const bool aligned = view_aligned(p->anchor_hm, p->W) && view_aligned(p->part_hm, p->W);
  if (is_f32 && aligned && p->W <= kPanelW * kMaxConsumers) return SDNET_PATH_CTA;
  return SDNET_PATH_WARP;
}

template <typename Kern>
void launch_peaks(Kern kern, dim3 grid, dim3 block, cudaStream_t stream, const PeaksParams& pp) {
  // > 48 KB of dynamic shared memory needs the opt-in (idempotent, cheap)
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeaksSmem);
  kern<<<grid, block, kPeaksSmem, stream>>>(pp);
}

// Launch with programmatic stream serialization (PDL): see pdl_wait() in the kernels.
template <typename Kern, typename Params>
void launch_pdl(Kern kern, dim3 grid, dim3 block, cudaStream_t stream, const Params& prm) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, prm);
}

int launch_decode(const SdnetDecodeParams* p, cudaStream_t stream, cudaEvent_t* marks = nullptr) {
  const Workspace ws = plan_workspace(p->B, p->M, p->N, p->H, p->W, p->K, p->P);
  char* base = static_cast<char*>(p->workspace);
  const int C = p->M + p->N;
  const size_t planes = (size_t)p->B * C;
  cudaError_t err = cudaMemsetAsync(base, 0, ws.off_lists, stream);
  if (err != cudaSuccess) return (int)err;
  if (marks) cudaEventRecord(marks[0], stream);

  PeaksParams pp;
  pp.anchor = to_view(p->anchor_hm);
  pp.part = to_view(p->part_hm);
  pp.B = p->B; pp.M = p->M; pp.N = p->N; pp.H = p->H; pp.W = p->W; pp.K = p->K; pp.P = p->P;
  pp.cap = ws.cap;
  pp.pre_activated = (p->flags & SDNET_FLAG_PRE_ACTIVATED) ? 1 : 0;
  pp.lists = reinterpret_cast<u64*>(base + ws.off_lists);
  pp.counts = reinterpret_cast<int*>(base + ws.off_counts);
  pp.sched = reinterpret_cast<u32*>(base + ws.off_sched);
  pp.ghist = reinterpret_cast<u32*>(base + ws.off_ghist);
  pp.gfloor = reinterpret_cast<int*>(base + ws.off_gfloor);
  pp.l2_prefetch_groups = 0;
  pp.tier1_units = 0;
  pp.tier1_planes = 0;
  const int sms = device_sm_count();
  CUtensorMap tm_anchor, tm_part;
  int tile_s = 0;
  const int path = select_peaks_path(p, &tm_anchor, &tm_part, &tile_s);
  const bool use_tile = path == SDNET_PATH_TILE || path == SDNET_PATH_TILE_ROW_PAIRS, use_cta = path == SDNET_PATH_CTA;
  const bool is_f32 = p->dtype == SDNET_DTYPE_F32;
  pp.odd_x = (int)(p->anchor_hm.stride_h / (is_f32 ? 1 : 2));  // in tensor-map elements
  static const int tier2_strips = [] {  // tuning knob, read once: SDNET_TIER2_STRIPS = n (default 2)
    const char* e = getenv("SDNET_TIER2_STRIPS");
    return e && atoi(e) > 0 ? atoi(e) : 2;
  }();
  static const int strips_override = [] {  // tuning knob, read once: SDNET_STRIPS = n
    const char* e = getenv("SDNET_STRIPS");
    return e ? atoi(e) : 0;
  }();
  auto pick_strips = [&](long long units_per_strip1, long long want_units) {
    int strips = (int)((want_units + units_per_strip1 - 1) / units_per_strip1);
    if (strips_override > 0) strips = strips_override;
    const int max_strips = (p->H + 31) / 32;  // strips of at least 32 rows
    if (strips > max_strips) strips = max_strips;
    if (strips < 1) strips = 1;
    pp.rows_per_strip = (p->H + strips - 1) / strips;
    if (use_tile && tile_s == 2) pp.rows_per_strip = (pp.rows_per_strip + 1) & ~1;  // row pairs: strips start on even rows
    pp.strips = (p->H + pp.rows_per_strip - 1) / pp.rows_per_strip;
  };
  if (use_tile) {
    const TileKernel kern = tile_kernel_for(p->radius, p->dtype, tile_s);
    const int kPanelCols = is_f32 ? TileGeom<SDNET_DTYPE_F32>::kPanel : TileGeom<SDNET_DTYPE_F16>::kPanel;
    const int kTileSmem = tile_smem(tile_s);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    static int per_sm_cache[3][3][3] = {};  // [dtype][rows per TMA row][radius]; 0 = not asked yet
    int& per_sm = per_sm_cache[p->dtype][tile_s][p->radius];
    if (per_sm == 0 &&
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileWarps * 32, kTileSmem) != cudaSuccess || per_sm < 1))
      per_sm = 1;
    pp.panels = (p->W + kPanelCols - 1) / kPanelCols;
    const long long resident_warps = (long long)sms * per_sm * kTileWarps;
    // long units prune best (measured: 128 images, 1 strip 0.169 ms, 7 strips 0.221 ms): split planes
    // into strips only while there are fewer units than resident warps
    {
      const long long units1 = (long long)planes * pp.panels;
      pp.tier1_units = 0;
      pp.tier1_planes = 0;
      if (units1 <= resident_warps || strips_override > 0) {
        // Fewer units than resident warps: every warp gets at most one unit per wave, so the kernel
        // lasts ceil(units / warps) unit-times.  Cutting planes into S strips shortens the unit but
        // costs ~15 % per extra strip (each strip re-warms its pruning floor); pick the S that
        // minimises waves(S) * (1 + 0.15 (S - 1)) / S.  Measured at 128 images: S = 1, 2, 3, 4 ->
        // 0.151, 0.163, 0.138, 0.183 ms.
        int best_s = 1;
        double best_cost = 1e30;
        const int max_s = (p->H + 31) / 32 < 16 ? (p->H + 31) / 32 : 16;
        for (int cand = 1; cand <= (max_s > 0 ? max_s : 1); ++cand) {
          const long long waves = (units1 * cand + resident_warps - 1) / resident_warps;
          const double cost = (double)waves * (1.0 + 0.15 * (cand - 1)) / cand;
          if (cost < best_cost - 1e-9) { best_cost = cost; best_s = cand; }
        }
        pick_strips(units1, units1 * best_s);
        pp.units = (int)(planes * pp.strips * pp.panels);
      } else {
        // whole waves of whole-height units first, then the leftover planes in short strips
        const long long full_waves = units1 / resident_warps;
        pp.tier1_planes = (int)((full_waves * resident_warps) / pp.panels);
        pp.tier1_units = pp.tier1_planes * pp.panels;
        const long long rest = (long long)planes - pp.tier1_planes;
        pick_strips(rest > 0 ? rest * pp.panels : 1, rest > 0 ? rest * pp.panels * tier2_strips : 1);
        pp.units = (int)(pp.tier1_units + rest * pp.strips * pp.panels);
      }
    }
    long long ctas = ((long long)pp.units + kTileWarps - 1) / kTileWarps;
    if (ctas > (long long)sms * per_sm) ctas = (long long)sms * per_sm;
    kern<<<dim3((unsigned)ctas), dim3(kTileWarps * 32), kTileSmem, stream>>>(pp, tm_anchor, tm_part);
  } else if (use_cta) {
    static const int ring_groups = [] {  // tuning knob, read once: SDNET_RING_GROUPS = 4 | 8
      const char* e = getenv("SDNET_RING_GROUPS");
      return (e && atoi(e) == 8) ? 8 : 4;
    }();
    static const int l2_pf = [] {
      const char* e = getenv("SDNET_L2_PREFETCH_GROUPS");
      return e ? atoi(e) : 0;
    }();
    pp.l2_prefetch_groups = l2_pf;
    const CtaGeom geom = cta_geometry(p->W, ring_groups);
    auto kern = ring_groups == 8
                    ? (p->radius == 2 ? sdnet_peaks_cta_kernel<2, 8> : sdnet_peaks_cta_kernel<1, 8>)
                    : (p->radius == 2 ? sdnet_peaks_cta_kernel<2, 4> : sdnet_peaks_cta_kernel<1, 4>);
    const int threads = (geom.nc + 1) * 32;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, geom.smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, geom.smem) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    pp.panels = 1;
    const long long resident = (long long)sms * per_sm;
    pick_strips((long long)planes, 3 * resident);
    pp.units = (int)(planes * pp.strips);
    long long ctas = pp.units < resident ? pp.units : resident;
    kern<<<dim3((unsigned)ctas), dim3(threads), geom.smem, stream>>>(pp);
  } else {
    pp.panels = (p->W + kPanelW - 1) / kPanelW;
    const int resident_warps = sms * kPeaksCtasPerSm * kWarps;
    pick_strips((long long)planes * pp.panels, 4ll * resident_warps);
    pp.units = (int)(planes * pp.strips * pp.panels);
    long long ctas = ((long long)pp.units + kWarps - 1) / kWarps;
    if (ctas > (long long)sms * kPeaksCtasPerSm) ctas = (long long)sms * kPeaksCtasPerSm;
    dim3 grid((unsigned)ctas), block(kThreads);
    const bool r2 = p->radius == 2;
    if (p->dtype == SDNET_DTYPE_F16) {
      if (r2) launch_peaks(sdnet_peaks_kernel<false, 2, SDNET_DTYPE_F16>, grid, block, stream, pp);
      else launch_peaks(sdnet_peaks_kernel<false, 1, SDNET_DTYPE_F16>, grid, block, stream, pp);
    } else if (p->dtype == SDNET_DTYPE_BF16) {
      if (r2) launch_peaks(sdnet_peaks_kernel<false, 2, SDNET_DTYPE_BF16>, grid, block, stream, pp);
      else launch_peaks(sdnet_peaks_kernel<false, 1, SDNET_DTYPE_BF16>, grid, block, stream, pp);
    } else {
      if (r2) launch_peaks(sdnet_peaks_kernel<false, 2, SDNET_DTYPE_F32>, grid, block, stream, pp);
      else launch_peaks(sdnet_peaks_kernel<false, 1, SDNET_DTYPE_F32>, grid, block, stream, pp);
    }
  }
  err = cudaGetLastError();
  if (err != cudaSuccess) return (int)err;
  if (marks) cudaEventRecord(marks[1], stream);

  {
    ExactParams ep;
    ep.anchor = pp.anchor; ep.part = pp.part;
    ep.B = p->B; ep.M = p->M; ep.N = p->N; ep.H = p->H; ep.W = p->W; ep.K = p->K; ep.P = p->P;
    ep.radius = p->radius; ep.cap = ws.cap;
    ep.force = (p->flags & SDNET_FLAG_EXACT_SELECT) ? 1 : 0;
    ep.pre_activated = pp.pre_activated;
    ep.lists = pp.lists; ep.counts = pp.counts;
    ep.flags = reinterpret_cast<int*>(base + ws.off_flags);
    {
      const size_t exact_grid = planes < (size_t)sms * 2 ? planes : (size_t)sms * 2;
      if (p->dtype == SDNET_DTYPE_F16) launch_pdl(sdnet_exact_select_kernel<SDNET_DTYPE_F16>, dim3((unsigned)exact_grid), dim3(kExactThreads), stream, ep);
      else if (p->dtype == SDNET_DTYPE_BF16) launch_pdl(sdnet_exact_select_kernel<SDNET_DTYPE_BF16>, dim3((unsigned)exact_grid), dim3(kExactThreads), stream, ep);
      else launch_pdl(sdnet_exact_select_kernel<SDNET_DTYPE_F32>, dim3((unsigned)exact_grid), dim3(kExactThreads), stream, ep);
    }
    err = cudaGetLastError();
    if (err != cudaSuccess) return (int)err;
    if (marks) cudaEventRecord(marks[2], stream);
  }

  TailParams tp;
  tp.offsets = to_view(p->offsets);
  tp.embeddings = to_view(p->embeddings);
  tp.B = p->B; tp.M = p->M; tp.N = p->N; tp.H = p->H; tp.W = p->W; tp.K = p->K; tp.P = p->P;
  tp.cap = ws.cap;
  tp.pre_activated = pp.pre_activated;
  tp.no_grouping = (p->flags & SDNET_FLAG_NO_GROUPING) ? 1 : 0;
  tp.conf = p->conf_f32;
  tp.dist_abs = p->dist_abs_f32;
  tp.lists = pp.lists;
  tp.counts = pp.counts;
  tp.anchor_out = p->anchor_out;
  tp.part_out = p->part_out;
  tp.anchor_inds = reinterpret_cast<long long*>(p->anchor_inds);
  tp.part_inds = reinterpret_cast<long long*>(p->part_inds);
  tp.part_emb = p->part_emb;
  tp.assign = p->assign;
  tp.out_counts = p->counts;
  tp.diag = p->diag;
  tp.exact_flags = reinterpret_cast<const int*>(base + ws.off_flags);
  tp.n_dest = p->n_dest;
  for (int j = 0; j < SDNET_MAX_DEST; ++j) tp.dest_delta[j] = j < p->n_dest ? p->dest_delta[j] : 0;
  if (p->dtype == SDNET_DTYPE_F16) launch_pdl(sdnet_tail_kernel<SDNET_DTYPE_F16>, dim3((unsigned)p->B), dim3(2 * kTeamThreads), stream, tp);
  else if (p->dtype == SDNET_DTYPE_BF16) launch_pdl(sdnet_tail_kernel<SDNET_DTYPE_BF16>, dim3((unsigned)p->B), dim3(2 * kTeamThreads), stream, tp);
  else launch_pdl(sdnet_tail_kernel<SDNET_DTYPE_F32>, dim3((unsigned)p->B), dim3(2 * kTeamThreads), stream, tp);
  err = cudaGetLastError();
  if (marks) cudaEventRecord(marks[3], stream);
  return (int)err;
}

}  // namespace

extern "C" {

int sdnet_abi_version(void) { return SDNET_ABI_VERSION; }

const char* sdnet_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case SDNET_E_NULL: return "a required pointer is NULL";
    case SDNET_E_SHAPE: return "bad shape (non-positive dim, H*W >= 2^24, k out of range, or above SDNET_MAX_*)";
    case SDNET_E_STRIDE: return "innermost stride must be 1";
    case SDNET_E_DTYPE: return "unsupported dtype";
    case SDNET_E_WORKSPACE: return "workspace missing, misaligned or too small";
    case SDNET_E_RADIUS: return "unsupported NMS radius";
    case SDNET_E_STRUCT: return "SdnetDecodeParams.struct_size mismatch";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

int sdnet_decode_workspace_bytes(int B, int M, int N, int H, int W, int K, int P, int dtype, size_t* out_bytes) {
  if (!out_bytes) return SDNET_E_NULL;
  if (dtype != SDNET_DTYPE_F32 && dtype != SDNET_DTYPE_F16 && dtype != SDNET_DTYPE_BF16) return SDNET_E_DTYPE;
  if (B <= 0 || M <= 0 || N <= 0 || H <= 0 || W <= 0 || K <= 0 || P <= 0) return SDNET_E_SHAPE;
  if ((long long)H * W >= (1ll << 24)) return SDNET_E_SHAPE;
  *out_bytes = plan_workspace(B, M, N, H, W, K, P).total;
  return 0;
}

int sdnet_match_launch(const SdnetMatchParams* p, void* stream) {
  if (!p) return SDNET_E_NULL;
  if (p->struct_size != sizeof(SdnetMatchParams)) return SDNET_E_STRUCT;
  if (!p->anchor_out || !p->part_out || !p->image_scale || !p->n_gt_anchors || !p->n_gt_parts || !p->anchor_stats ||
      !p->part_stats || !p->anchor_acc || !p->part_acc || (p->max_gt_anchors > 0 && !p->gt_anchors) ||
      (p->max_gt_parts > 0 && !p->gt_parts))
    return SDNET_E_NULL;
  if (p->B <= 0 || p->M <= 0 || p->N <= 0 || p->K <= 0 || p->P <= 0 || p->K > SDNET_MAX_TOPK || p->P > SDNET_MAX_TOPK ||
      p->M + p->N > SDNET_MAX_CHANNELS || p->max_gt_anchors < 0 || p->max_gt_parts < 0 || p->max_gt_anchors > SDNET_MAX_GT ||
      p->max_gt_parts > SDNET_MAX_GT)
    return SDNET_E_SHAPE;
  sdnet_match_kernel<<<dim3((unsigned)p->B), dim3(kMatchThreads), 0, static_cast<cudaStream_t>(stream)>>>(*p);
  return (int)cudaGetLastError();
}

int sdnet_decode_peaks_path(const SdnetDecodeParams* params) {
  const int rc = validate(params);
  if (rc != 0) return rc;
  CUtensorMap a, b;
  int tile_s = 0;
  return select_peaks_path(params, &a, &b, &tile_s);
}

int sdnet_decode_launch(const SdnetDecodeParams* params, void* stream) {
  const int rc = validate(params);
  if (rc) return rc;
  return launch_decode(params, static_cast<cudaStream_t>(stream));
}

int sdnet_decode_launch_timed(const SdnetDecodeParams* params, void* stream_v, float* kernel_ms) {
  const int rc = validate(params);
  if (rc) return rc;
  if (!kernel_ms) return SDNET_E_NULL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  cudaEvent_t marks[4];
  for (int i = 0; i < 4; ++i) {
    cudaError_t e = cudaEventCreate(&marks[i]);
    if (e != cudaSuccess) return (int)e;
  }
  int out = launch_decode(params, stream, marks);
  if (out == 0) {
    cudaError_t e = cudaEventSynchronize(marks[3]);
    if (e != cudaSuccess) out = (int)e;
    for (int i = 0; i < 3 && out == 0; ++i) {
      e = cudaEventElapsedTime(&kernel_ms[i], marks[i], marks[i + 1]);
      if (e != cudaSuccess) out = (int)e;
    }
  }
  for (int i = 0; i < 4; ++i) cudaEventDestroy(marks[i]);
  return out;
}

int sdnet_activate_launch(const SdnetTensor4* in, int dtype, int B, int C, int H, int W, float* out, void* stream) {
  if (!in || !in->data || !out) return SDNET_E_NULL;
  if (dtype != SDNET_DTYPE_F32 && dtype != SDNET_DTYPE_F16 && dtype != SDNET_DTYPE_BF16) return SDNET_E_DTYPE;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return SDNET_E_SHAPE;
  if (in->stride_w != 1) return SDNET_E_STRIDE;
  const size_t total = (size_t)B * C * H * W;
  const int sms = device_sm_count();
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)sms * 8) blocks = (size_t)sms * 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SDNET_DTYPE_F16) sdnet_activate_kernel<SDNET_DTYPE_F16><<<dim3((unsigned)blocks), dim3(256), 0, st>>>(to_view(*in), C, H, W, total, out);
  else if (dtype == SDNET_DTYPE_BF16) sdnet_activate_kernel<SDNET_DTYPE_BF16><<<dim3((unsigned)blocks), dim3(256), 0, st>>>(to_view(*in), C, H, W, total, out);
  else sdnet_activate_kernel<SDNET_DTYPE_F32><<<dim3((unsigned)blocks), dim3(256), 0, st>>>(to_view(*in), C, H, W, total, out);
  return (int)cudaGetLastError();
}

int sdnet_decode_host_launch(const SdnetDecodeParams* params, void* staging, size_t staging_bytes, void* stream_v) {
  const int rc = validate(params);
  if (rc) return rc;
  if (!staging) return SDNET_E_NULL;
  if (params->dtype != SDNET_DTYPE_F32) return SDNET_E_DTYPE;  // host-buffer entry point: fp32 only for now
  const SdnetDecodeParams& p = *params;
  const size_t plane_bytes = (size_t)p.H * p.W * sizeof(float);
  const size_t need = (size_t)p.B * (p.M + p.N) * plane_bytes;
  if (staging_bytes < need || ((uintptr_t)staging & 255)) return SDNET_E_WORKSPACE;
  if (p.anchor_hm.stride_h != p.W || p.part_hm.stride_h != p.W) return SDNET_E_STRIDE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  // heat planes: host (strided per image) -> dense device staging [B][M+N][H][W]
  char* dst = static_cast<char*>(staging);
  const size_t img_bytes = (size_t)(p.M + p.N) * plane_bytes;
  cudaError_t err;
  if (p.anchor_hm.stride_c == (long long)p.H * p.W) {
    err = cudaMemcpy2DAsync(dst, img_bytes, p.anchor_hm.data, (size_t)p.anchor_hm.stride_b * sizeof(float),
                            (size_t)p.M * plane_bytes, p.B, cudaMemcpyHostToDevice, stream);
  } else {
    err = cudaErrorInvalidValue;
  }
  if (err != cudaSuccess) return (int)err;
  if (p.part_hm.stride_c == (long long)p.H * p.W || p.N == 1) {
    err = cudaMemcpy2DAsync(dst + (size_t)p.M * plane_bytes, img_bytes, p.part_hm.data,
                            (size_t)p.part_hm.stride_b * sizeof(float), (size_t)p.N * plane_bytes, p.B,
                            cudaMemcpyHostToDevice, stream);
  } else {
    err = cudaErrorInvalidValue;
  }
  if (err != cudaSuccess) return (int)err;
  SdnetDecodeParams q = p;
  q.anchor_hm.data = dst;
  q.anchor_hm.stride_b = (long long)(p.M + p.N) * p.H * p.W;
  q.anchor_hm.stride_c = (long long)p.H * p.W;
  q.part_hm.data = dst + (size_t)p.M * plane_bytes;
  q.part_hm.stride_b = q.anchor_hm.stride_b;
  q.part_hm.stride_c = q.anchor_hm.stride_c;
  // offsets / embeddings stay in pinned host memory: the tail kernel reads them through
  // the unified address space only at the K + 2P selected pixels per image.
  return launch_decode(&q, stream);
}

}  // extern "C"
