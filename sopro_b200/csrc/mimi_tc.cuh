// tcgen05 / TMEM / TMA implicit GEMM for the Mimi decoder's dense blocks (sm_100a only).
//
//   C[b][m][n] = epi( sum_{j<taps} sum_{ci<Cin} X[b][m + j*dil - pad][ci] * W[n][j*Cin + ci] + bias[n % bias_mod] )
//
// X is a channel-last bf16 activation [B][Min][Cin] (already passed through ELU by its producer when
// the layer wants ELU(x)), W is a bf16 weight matrix [N][K] (K = taps*Cin, K-major), accumulation is
// fp32 in tensor memory.  This covers Linear, the causal Conv1d and the causal ConvTranspose1d of
// transformers' modeling_mimi.py (:331-351, :402-409) exactly as mimi_engine.cu's fp32 path does.
//
// One CTA computes one 128 x BN tile:
//   warp 0    TMA producer: per 64-wide K chunk one 3-D box of X (rows shifted by the tap; rows outside
//             [0, Min) are zero-filled by the TMA unit == the causal left pad) and one 2-D box of W,
//             both landing 128B-swizzled in a ring of shared-memory stages, completion on mbarriers
//   warp 1    allocates BN tensor-memory columns, one lane issues tcgen05.mma (M=128, N=BN, K=16) four
//             times per chunk; tcgen05.commit releases the stage / signals the accumulator
//   warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns, bias / GELU / LayerScale-residual / skip,
//             writes fp32 and/or bf16 (optionally ELU'd: the next layer's operand)
// Two CTAs are resident per SM (<= 100 KB of stages each), so one tile's epilogue overlaps the other's
// main loop.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>

namespace tc {

enum { EPI_NONE = 0, EPI_GELU = 1, EPI_RES_SCALE = 2, EPI_RES = 3 };

struct TcOp {
  const float* bias;   // [bias_mod] or null
  const float* R;      // residual fp32 [B][M][N] or null
  const float* scale;  // LayerScale [N] or null
  float* out_f32;      // [B][M][N] or null
  __nv_bfloat16* out_bf16;  // [B][M][N] or null
  long long c_bs;      // batch stride of R / outputs (elements)
  int M, N, K, Cin, dil, pad, bias_mod, epi, out_elu;
  int stages;          // filled in by launch()
  int tap_col[8];      // A column offset of tap j (0 everywhere for convolutions; the NAR refiner's exact three-way bf16 split
                       // pairs W's K block j with the A term it multiplies: taps over COLUMN groups of the same rows, dil = 0)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra W_%=;\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> one row slice per thread
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, K-major operand, 128-byte swizzle, rows of 64 bf16 (128 B) stacked
// densely: 8-row groups are 1024 B apart (SBO), one swizzle atom along K (LBO unused).
// Bit layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor): addr>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// same for rows of 32 bf16 (64 B), 64-byte swizzle: 8-row groups 512 B apart, layout_type SWIZZLE_64B = 4
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
template <int BK>
__device__ __forceinline__ uint64_t smem_desc_k(uint32_t addr) {
  return BK == 64 ? smem_desc_sw128(addr) : smem_desc_sw64(addr);
}
// instruction descriptor, kind::f16: D fp32 (bits [4,6) = 1), A/B bf16 ([7,10) = [10,13) = 1), both
// K-major (bits 15, 16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// 256-bit global accesses (one full 32-byte sector per thread per instruction)
__device__ __forceinline__ void stg256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
// ELU whose result is rounded to bf16 right away: exp(x) - 1 with the fast exponential is exact to well below
// half a bf16 ulp (absolute error ~1e-7 against a result of magnitude >= |x|/2)
__device__ __forceinline__ float elu_fast(float x) { return x > 0.f ? x : __expf(x) - 1.0f; }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

constexpr int kBM = 128, kBK = 64, kThreads = 192;
constexpr int kGemmThreads = 320;  // TMA warp, MMA warp, 8 epilogue warps (2 per tensor-memory lane quarter)
template <int BN, int BK>
struct TileCfg {
  static constexpr int kMaxStages = BN >= 128 ? 3 : 4;
  static constexpr int kABytes = kBM * BK * 2;  // 16 KB at BK = 64
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int smem(int stages) { return stages * kStageBytes + 1024; }  // + alignment slack
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
};

template <int BN, int BK>
__global__ void __launch_bounds__(kGemmThreads, 3) igemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmW, const TcOp op) {
  using Cfg = TileCfg<BN, BK>;
  constexpr int SM = Cfg::kMaxStages;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * SM + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[SM]), accbar = smem_u32(&bars[2 * SM]);
  const int S = op.stages;  // min(kMaxStages, K chunks): short-K layers take less shared memory -> more CTAs per SM
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN, b = blockIdx.z;
  const int nk = op.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % S;
        const uint32_t ph = (uint32_t)(kc / S) & 1u;
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        mbar_expect_tx(full0 + 8 * s, Cfg::kStageBytes);
        const int k0 = kc * BK;
        const int j = k0 / op.Cin, ci = k0 - j * op.Cin;
        const uint32_t sa = tiles + s * Cfg::kStageBytes;
        tma_load_3d(sa, &tmA, full0 + 8 * s, ci + op.tap_col[j & 7], m0 + j * op.dil - op.pad, b);
        tma_load_2d(sa + Cfg::kABytes, &tmW, full0 + 8 * s, k0, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(kBM, BN);
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % S;
        const uint32_t ph = (uint32_t)(kc / S) & 1u;
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t sa = tiles + s * Cfg::kStageBytes;
        const uint64_t da = smem_desc_k<BK>(sa), db = smem_desc_k<BK>(sa + Cfg::kABytes);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)  // +32 B per K=16 slice inside the swizzle atom
          tc_mma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0);
        tc_commit(empty0 + 8 * s);
      }
      tc_commit(accbar);
    }
  } else {
    mbar_wait(accbar, 0);
    tc_fence_after();
    const int q = warp & 3;  // tensor-memory lane quarter this warp may read
    const int m = m0 + 32 * q + lane;
    const bool row_ok = m < op.M;
    const size_t row = (size_t)b * (size_t)op.c_bs + (size_t)m * op.N;
    // the two warps of a quarter split the columns (a 32-column tile is done by the first alone)
    constexpr int kColsPerWarp = BN >= 64 ? BN / 2 : BN;
    const int cbeg = BN >= 64 ? ((warp - 2) >> 2) * kColsPerWarp : ((warp - 2) >> 2) * BN;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + kColsPerWarp && c0 < BN; c0 += 32) {
      uint32_t r[32];
      tc_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, r);
      if (!row_ok) continue;
      const int n = n0 + c0;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      if (op.bias) {
        const int nb = n % op.bias_mod;  // bias_mod is a multiple of 32 whenever N > bias_mod
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(op.bias + nb + i));
          v[i] += bv.x;
          v[i + 1] += bv.y;
          v[i + 2] += bv.z;
          v[i + 3] += bv.w;
        }
      }
      if (op.epi == EPI_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
      } else if (op.epi == EPI_RES_SCALE || op.epi == EPI_RES) {
        const float* rp = op.R + row + n;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float rv[8];
          ldg256(rp + i, rv);
          if (op.epi == EPI_RES_SCALE) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(op.scale + n + i));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(op.scale + n + i + 4));
            const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) v[i + e] = fmaf(sv[e], v[i + e], rv[e]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[i + e] += rv[e];
          }
        }
      }
      if (op.out_f32) {
        float* o = op.out_f32 + row + n;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint32_t w8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w8[e] = __float_as_uint(v[i + e]);
          stg256(o + i, w8);
        }
      }
      if (op.out_bf16) {
        __nv_bfloat16* o = op.out_bf16 + row + n;
#pragma unroll
        for (int i = 0; i < 32; i += 16) {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a = v[i + 2 * e], c = v[i + 2 * e + 1];
            if (op.out_elu) {
              a = elu_fast(a);
              c = elu_fast(c);
            }
            const __nv_bfloat162 h = __floats2bfloat162_rn(a, c);
            pk[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          stg256(o + i, pk);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Fused ResnetBlock (MimiResnetBlock, modeling_mimi.py:437-451): out = z + conv1x1(ELU(conv3(ELU(z)))).
//   GEMM 1: h[128 x HID] = conv k=3 of the bf16 ELU(z) rows (3-D TMA boxes per tap), accumulators in tensor
//           memory columns [0, HID)
//   epilogue A: + bias, ELU, round to bf16, written swizzled into shared memory as the next A operand
//   GEMM 2: [128 x 2*HID] = h . W2^T (W2 loaded once per CTA) into columns [HID, 3*HID)
//   epilogue B: + bias + z (fp32 skip, read once), then fp32 and/or bf16(ELU) out
// The hidden activation never touches HBM and one launch replaces two.  HID in {32, 64, 128}.
// ---------------------------------------------------------------------------------------------
struct ResOp {
  const float* bias1;  // [HID]
  const float* bias2;  // [2*HID]
  const float* Z;      // fp32 skip [B][M][2*HID]
  float* out_f32;      // [B][M][2*HID] or null
  __nv_bfloat16* out_bf16;  // [B][M][2*HID] or null, through ELU when out_elu
  int M, taps, pad, out_elu, stages;
  int Min;  // rows of the bf16 input X (0 = M; streaming: M + taps - 1 with pad = 0, the context rows in front)
};

template <int HID>
struct ResCfg {
  static constexpr int kCout = 2 * HID;
  static constexpr int kBKH = HID < 64 ? HID : 64;       // K chunk of the second GEMM
  static constexpr int kNK2 = HID / kBKH;
  static constexpr int kMaxStages = 3;
  static constexpr int kABytes = kBM * kBK * 2;           // 16 KB
  static constexpr int kBBytes = HID * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kHBytes = kBM * HID * 2;           // hidden activation as the A operand of GEMM 2
  static constexpr int kW2Bytes = kCout * HID * 2;
  static constexpr int kTmemCols = 3 * HID <= 128 ? 128 : (3 * HID <= 256 ? 256 : 512);
  static constexpr int smem(int stages) { return stages * kStageBytes + kHBytes + kW2Bytes + 1024; }
};

template <int HID>
__global__ void __launch_bounds__(kGemmThreads, HID <= 32 ? 3 : (HID <= 64 ? 2 : 1)) resblock_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                                     const __grid_constant__ CUtensorMap tmW1,
                                                                                     const __grid_constant__ CUtensorMap tmW2,
                                                                                     const ResOp op) {
  using Cfg = ResCfg<HID>;
  constexpr int SM = Cfg::kMaxStages, COUT = Cfg::kCout, BKH = Cfg::kBKH;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * SM + 4];
  __shared__ uint32_t tmem_slot;
  const int S = op.stages;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sH = tiles + S * Cfg::kStageBytes, sW2 = sH + Cfg::kHBytes;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[SM]);
  const uint32_t w2_full = smem_u32(&bars[2 * SM]), acc1 = smem_u32(&bars[2 * SM + 1]), h_ready = smem_u32(&bars[2 * SM + 2]),
                 acc2 = smem_u32(&bars[2 * SM + 3]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM, b = blockIdx.y;
  const int nk = op.taps * COUT / kBK;  // K chunks of GEMM 1 (COUT is a multiple of 64)

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(w2_full, 1);
    mbar_init(acc1, 1);
    mbar_init(h_ready, kGemmThreads - 64);
    mbar_init(acc2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w2_full, Cfg::kW2Bytes);
      for (int c = 0; c < Cfg::kNK2; ++c) tma_load_2d(sW2 + c * (COUT * BKH * 2), &tmW2, w2_full, c * BKH, 0);
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % S;
        const uint32_t ph = (uint32_t)(kc / S) & 1u;
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        mbar_expect_tx(full0 + 8 * s, Cfg::kStageBytes);
        const int k0 = kc * kBK;
        const int j = k0 / COUT, ci = k0 - j * COUT;
        const uint32_t sa = tiles + s * Cfg::kStageBytes;
        tma_load_3d(sa, &tmA, full0 + 8 * s, ci, m0 + j - op.pad, b);
        tma_load_2d(sa + Cfg::kABytes, &tmW1, full0 + 8 * s, k0, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % S;
        const uint32_t ph = (uint32_t)(kc / S) & 1u;
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t sa = tiles + s * Cfg::kStageBytes;
        const uint64_t da = smem_desc_sw128(sa), db = smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) tc_mma_bf16(tmem, da + 2 * k, db + 2 * k, instr_desc_bf16(kBM, HID), (kc | k) != 0);
        tc_commit(empty0 + 8 * s);
      }
      tc_commit(acc1);
      mbar_wait(w2_full, 0);
      mbar_wait(h_ready, 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < Cfg::kNK2; ++c) {
        const uint64_t dh = smem_desc_k<BKH>(sH + c * (kBM * BKH * 2)), dw = smem_desc_k<BKH>(sW2 + c * (COUT * BKH * 2));
#pragma unroll
        for (int k = 0; k < BKH / 16; ++k) tc_mma_bf16(tmem + HID, dh + 2 * k, dw + 2 * k, instr_desc_bf16(kBM, COUT), (c | k) != 0);
      }
      tc_commit(acc2);
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = 32 * q + lane;  // tile row == tensor-memory lane
    const int m = m0 + r;
    const bool row_ok = m < op.M;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    // ---- epilogue A: hidden activation -> shared memory (bf16, swizzled K-major rows of BKH elements)
    mbar_wait(acc1, 0);
    tc_fence_after();
    constexpr int kColsA = HID >= 64 ? HID / 2 : HID;
    const int abeg = HID >= 64 ? half * kColsA : half * HID;
#pragma unroll 1
    for (int c0 = abeg; c0 < abeg + kColsA && c0 < HID; c0 += 32) {
      uint32_t v[32];
      tc_ld32(trow + (uint32_t)c0, v);
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(op.bias1 + c0 + e));
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(elu_fast(__uint_as_float(v[e]) + bv.x), elu_fast(__uint_as_float(v[e + 1]) + bv.y));
        const __nv_bfloat162 h1 = __floats2bfloat162_rn(elu_fast(__uint_as_float(v[e + 2]) + bv.z), elu_fast(__uint_as_float(v[e + 3]) + bv.w));
        pk[e >> 1] = *reinterpret_cast<const uint32_t*>(&h0);
        pk[(e >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h1);
      }
      // 32 columns = 4 sixteen-byte chunks of row r inside K chunk c0 / BKH; chunk index XOR row bits (swizzle)
      const uint32_t blk = sH + (uint32_t)(c0 / BKH) * (kBM * BKH * 2) + (uint32_t)r * (BKH * 2);
      const int ch0 = (c0 % BKH) >> 3;
      const int sw = BKH == 64 ? (r & 7) : ((r >> 1) & 3);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const uint32_t a = blk + (uint32_t)(((ch0 + cc) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * cc]), "r"(pk[4 * cc + 1]), "r"(pk[4 * cc + 2]),
                     "r"(pk[4 * cc + 3])
                     : "memory");
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(h_ready) : "memory");
    // ---- epilogue B: + bias + skip, outputs
    mbar_wait(acc2, 0);
    tc_fence_after();
    const size_t row = ((size_t)b * (size_t)op.M + (size_t)(row_ok ? m : 0)) * COUT;
    constexpr int kColsB = COUT / 2;
#pragma unroll 1
    for (int c0 = half * kColsB; c0 < (half + 1) * kColsB; c0 += 32) {
      uint32_t rr[32];
      tc_ld32(trow + (uint32_t)(HID + c0), rr);
      if (!row_ok) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        float zv[8];
        ldg256(op.Z + row + c0 + i, zv);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(op.bias2 + c0 + i));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(op.bias2 + c0 + i + 4));
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i + e] = zv[e] + (__uint_as_float(rr[i + e]) + bb[e]);
      }
      if (op.out_f32) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint32_t w8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w8[e] = __float_as_uint(v[i + e]);
          stg256(op.out_f32 + row + c0 + i, w8);
        }
      }
      if (op.out_bf16) {
#pragma unroll
        for (int i = 0; i < 32; i += 16) {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a = v[i + 2 * e], c = v[i + 2 * e + 1];
            if (op.out_elu) {
              a = elu_fast(a);
              c = elu_fast(c);
            }
            const __nv_bfloat162 hh = __floats2bfloat162_rn(a, c);
            pk[e] = *reinterpret_cast<const uint32_t*>(&hh);
          }
          stg256(op.out_bf16 + row + c0 + i, pk);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Causal sliding-window attention on the tensor cores (MimiAttention.forward, modeling_mimi.py:681-738,
// head_dim 64, window <= 257).  One CTA = 128 queries of one (batch, head):
//   S = Q K^T over the 384 keys [q0-256, q0+128) -> 384 fp32 columns of tensor memory (3 x M128 N128 K64)
//   softmax with thread == query row (no shuffles): two passes over tensor memory, P rounded to bf16 and
//   written 128B-swizzled into shared memory as the next A operand (over the dead Q/K tiles)
//   O = P V with V^T [d][key] tiles as the K-major B operand -> 64 more columns; scaled by 1/sum on the way out
// Inputs are the rotated bf16 q / k [B][T2][C] and v transposed [B][C][T2p] written by rope_pack_kernel.
// ---------------------------------------------------------------------------------------------
struct AttnOp {
  __nv_bfloat16* out;  // [B][T2][C]
  int T2, C, window;
  float scale_log2e;   // log2(e) / sqrt(head_dim)
};

constexpr int kAttnKeys = 384, kAttnDh = 64;
constexpr int kAttnSmem = 65536 + 49152 + 32768 + 1024;

static __global__ void __launch_bounds__(kThreads) attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                           const __grid_constant__ CUtensorMap tmVt, const AttnOp op) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[5];
  __shared__ uint32_t tmem_slot;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = tiles, sK = tiles + 16384, sV = tiles + 65536;
  const uint32_t qk_full = smem_u32(&bars[0]), v_full = smem_u32(&bars[1]), s_full = smem_u32(&bars[2]),
                 p_full = smem_u32(&bars[3]), o_full = smem_u32(&bars[4]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int kbase = q0 + 128 - kAttnKeys;
  auto p_block = [&](int kb) -> uint32_t { return kb < 4 ? tiles + kb * 16384 : tiles + 65536 + 49152 + (kb - 4) * 16384; };

  if (threadIdx.x == 0) {
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(qk_full, 65536);
      tma_load_3d(sQ, &tmQ, qk_full, h * kAttnDh, q0, b);
      for (int kt = 0; kt < 3; ++kt) tma_load_3d(sK + kt * 16384, &tmK, qk_full, h * kAttnDh, kbase + kt * 128, b);
      mbar_expect_tx(v_full, 49152);
      for (int kb = 0; kb < 6; ++kb) tma_load_3d(sV + kb * 8192, &tmVt, v_full, kbase + kb * 64, h * kAttnDh, b);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(qk_full, 0);
      tc_fence_after();
      const uint64_t dq = smem_desc_sw128(sQ);
      for (int kt = 0; kt < 3; ++kt) {
        const uint64_t dk = smem_desc_sw128(sK + kt * 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem + kt * 128, dq + 2 * k, dk + 2 * k, instr_desc_bf16(128, 128), k != 0);
      }
      tc_commit(s_full);
      mbar_wait(p_full, 0);
      mbar_wait(v_full, 0);
      tc_fence_after();
      for (int kb = 0; kb < 6; ++kb) {
        const uint64_t dp = smem_desc_sw128(p_block(kb)), dv = smem_desc_sw128(sV + kb * 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem + kAttnKeys, dp + 2 * k, dv + 2 * k, instr_desc_bf16(128, 64), (kb | k) != 0);
      }
      tc_commit(o_full);
    }
  } else {
    const int q = warp & 3;
    const int r = 32 * q + lane;   // row of the tile == tensor-memory lane
    const int i = q0 + r;          // query position
    const bool row_ok = i < op.T2;
    // valid key columns of this row / of the whole warp (slices outside the warp's band are skipped uniformly)
    const int clo = row_ok ? max(0, i - op.window + 1) - kbase : 1, chi = row_ok ? i - kbase : 0;
    const int i_first = q0 + 32 * q, i_last = min(i_first + 31, op.T2 - 1);
    const int wlo = max(0, i_first - op.window + 1) - kbase, whi = i_last - kbase;  // whi < wlo when the warp has no rows
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    mbar_wait(s_full, 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < kAttnKeys; c0 += 32) {
      if (c0 + 31 < wlo || c0 > whi) continue;
      uint32_t v[32];
      tc_ld32(trow + c0, v);
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (c0 + e >= clo && c0 + e <= chi) mx = fmaxf(mx, __uint_as_float(v[e]));
    }
    float sum = 0.f;
    const float mxs = mx * op.scale_log2e;
#pragma unroll 1
    for (int c0 = 0; c0 < kAttnKeys; c0 += 32) {
      const bool live = !(c0 + 31 < wlo || c0 > whi);
      uint32_t pk[16];
      if (live) {
        uint32_t v[32];
        tc_ld32(trow + c0, v);
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float p0 = 0.f, p1 = 0.f;
          if (c0 + e >= clo && c0 + e <= chi) p0 = exp2f(fmaf(__uint_as_float(v[e]), op.scale_log2e, -mxs));
          if (c0 + e + 1 >= clo && c0 + e + 1 <= chi) p1 = exp2f(fmaf(__uint_as_float(v[e + 1]), op.scale_log2e, -mxs));
          const __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
          sum += __low2float(hh) + __high2float(hh);
          pk[e >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = 0u;
      }
      // 32 keys = 4 sixteen-byte chunks of row r in P block c0/64, chunk index XOR (r & 7)
      const uint32_t rowaddr = p_block(c0 >> 6) + r * 128;
      const int ch0 = (c0 & 63) >> 3;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const uint32_t a = rowaddr + (uint32_t)(((ch0 + cc) ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * cc]), "r"(pk[4 * cc + 1]), "r"(pk[4 * cc + 2]),
                     "r"(pk[4 * cc + 3])
                     : "memory");
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(p_full) : "memory");
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = row_ok ? 1.0f / sum : 0.f;
    __nv_bfloat16* orow = op.out + ((size_t)b * op.T2 + (row_ok ? i : 0)) * op.C + h * kAttnDh;
#pragma unroll 1
    for (int c0 = 0; c0 < kAttnDh; c0 += 32) {
      uint32_t v[32];
      tc_ld32(trow + kAttnKeys + c0, v);
      if (!row_ok) continue;
#pragma unroll
      for (int e = 0; e < 32; e += 8) {
        uint32_t w4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(v[e + 2 * u]) * inv, __uint_as_float(v[e + 2 * u + 1]) * inv);
          w4[u] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(orow + c0 + e) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// activations [B][rows][cin] bf16, box = bk channels x 128 rows x 1 batch
inline bool make_act_map(CUtensorMap* tm, const void* base, int B, long long rows, int cin, int bk = kBK) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)rows, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)rows * (cuuint64_t)cin * 2};
  const cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)kBM, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// weights [N][K] bf16, box = bk x BN
inline bool make_weight_map(CUtensorMap* tm, const void* base, int N, int K, int BN, int bk = kBK) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)BN};
  const cuuint32_t es[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// v transposed [B][C][T2p] bf16, box = 64 keys x 64 channels
inline bool make_vt_map(CUtensorMap* tm, const void* base, int B, int C, long long T2, long long T2p) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)T2, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)T2p * 2, (cuuint64_t)C * (cuuint64_t)T2p * 2};
  const cuuint32_t box[3] = {64, 64, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool attn_supported(int C, int H, int window) { return H > 0 && C == H * kAttnDh && window >= 1 && window <= kAttnKeys - 127; }

// q, k: bf16 [B][T2][C] (rotated); vt: bf16 [B][C][T2p]; out: bf16 [B][T2][C]
// the dynamic shared-memory opt-in is a per-device function attribute: remember it per device
inline bool attr_needed(unsigned long long& done_mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (done_mask >> dev & 1ull) return false;
  done_mask |= 1ull << dev;
  return true;
}

inline cudaError_t launch_attn(const void* q, const void* k, const void* vt, __nv_bfloat16* out, int B, int T2, long long T2p, int C,
                               int H, int window, cudaStream_t st) {
  static unsigned long long attr_done = 0;
  if (attr_needed(attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return e;
  }
  CUtensorMap tmQ, tmK, tmV;
  if (!make_act_map(&tmQ, q, B, T2, C) || !make_act_map(&tmK, k, B, T2, C) || !make_vt_map(&tmV, vt, B, C, T2, T2p))
    return cudaErrorInvalidValue;
  AttnOp op{out, T2, C, window, 1.4426950408889634f / sqrtf((float)kAttnDh)};
  dim3 grid((unsigned)((T2 + 127) / 128), (unsigned)H, (unsigned)B);
  attn_tc_kernel<<<grid, kThreads, kAttnSmem, st>>>(tmQ, tmK, tmV, op);
  return cudaGetLastError();
}

inline int pick_bn(int N) { return N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : (N % 32 == 0 ? 32 : 0)); }

// ---- fused ResnetBlock launcher.  X: bf16 ELU(z) [B][M][2*hid]; W1: bf16 [hid][taps*2*hid]; W2: bf16 [2*hid][hid]
inline bool resblock_supported(int hid, int cout) { return cout == 2 * hid && (hid == 32 || hid == 64 || hid == 128); }

template <int HID>
inline cudaError_t launch_resblock_t(const void* X, const void* W1, const void* W2, ResOp op, int B, cudaStream_t st) {
  using Cfg = ResCfg<HID>;
  static unsigned long long attr_done = 0;
  if (attr_needed(attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(resblock_tc_kernel<HID>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem(Cfg::kMaxStages));
    if (e != cudaSuccess) return e;
  }
  const int cout = Cfg::kCout, K1 = op.taps * cout, nk = K1 / kBK;
  op.stages = nk < Cfg::kMaxStages ? nk : Cfg::kMaxStages;
  CUtensorMap tmA, tmW1, tmW2;
  if (!make_act_map(&tmA, X, B, op.Min > 0 ? op.Min : op.M, cout) || !make_weight_map(&tmW1, W1, HID, K1, HID) ||
      !make_weight_map(&tmW2, W2, cout, HID, cout, Cfg::kBKH))
    return cudaErrorInvalidValue;
  dim3 grid((unsigned)((op.M + kBM - 1) / kBM), (unsigned)B);
  resblock_tc_kernel<HID><<<grid, kGemmThreads, Cfg::smem(op.stages), st>>>(tmA, tmW1, tmW2, op);
  return cudaGetLastError();
}

inline cudaError_t launch_resblock(const void* X, const void* W1, const void* W2, int hid, const ResOp& op, int B, cudaStream_t st) {
  switch (hid) {
    case 32: return launch_resblock_t<32>(X, W1, W2, op, B, st);
    case 64: return launch_resblock_t<64>(X, W1, W2, op, B, st);
    default: return launch_resblock_t<128>(X, W1, W2, op, B, st);
  }
}

// K chunk width: 64 channels (128-byte rows) when the channel count allows, else 32 (64-byte rows)
inline int pick_bk(int Cin) { return Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 0); }

// supported when every K chunk stays inside one tap and the tile shapes divide
inline bool supported(int N, int K, int Cin) { return pick_bk(Cin) != 0 && K % Cin == 0 && pick_bn(N) != 0; }

template <int BN, int BK>
inline cudaError_t launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmW, const TcOp& op, int B, cudaStream_t st) {
  using Cfg = TileCfg<BN, BK>;
  static unsigned long long attr_done = 0;
  if (attr_needed(attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(igemm_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem(Cfg::kMaxStages));
    if (e != cudaSuccess) return e;
  }
  TcOp o = op;
  const int nk = op.K / BK;
  o.stages = nk < Cfg::kMaxStages ? nk : Cfg::kMaxStages;
  // short-K layers are bound by their epilogues: two stages leave room for a third resident CTA per SM
  if (nk <= 4 && o.stages > 2) o.stages = 2;
  dim3 grid((unsigned)((op.M + kBM - 1) / kBM), (unsigned)(op.N / BN), (unsigned)B);
  igemm_tc_kernel<BN, BK><<<grid, kGemmThreads, Cfg::smem(o.stages), st>>>(tmA, tmW, o);
  return cudaGetLastError();
}

// X: bf16 [B][Min][a_cols] (a_cols = 0: Cin columns; larger when op.tap_col addresses column groups); W: bf16 [N][K]
inline cudaError_t launch(const void* X, long long Min, const void* W, const TcOp& op, int B, cudaStream_t st, int a_cols = 0) {
  const int BN = pick_bn(op.N), BK = pick_bk(op.Cin);
  CUtensorMap tmA, tmW;
  if (!make_act_map(&tmA, X, B, Min, a_cols > 0 ? a_cols : op.Cin, BK) || !make_weight_map(&tmW, W, op.N, op.K, BN, BK))
    return cudaErrorInvalidValue;
  if (BK == 64) {
    switch (BN) {
      case 128: return launch_bn<128, 64>(tmA, tmW, op, B, st);
      case 64: return launch_bn<64, 64>(tmA, tmW, op, B, st);
      default: return launch_bn<32, 64>(tmA, tmW, op, B, st);
    }
  }
  switch (BN) {  // narrow layers (Cin = 32)
    case 128: return launch_bn<128, 32>(tmA, tmW, op, B, st);
    case 64: return launch_bn<64, 32>(tmA, tmW, op, B, st);
    default: return launch_bn<32, 32>(tmA, tmW, op, B, st);
  }
}

}  // namespace tc
