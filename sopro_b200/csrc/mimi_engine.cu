// Mimi codec DECODE path on sm_100a: codes [B, Q, T] -> wav [B, T*1920].
//
// Replaces transformers.MimiModel.decode as called by the reference (codec/mimi.py:65-72):
// RVQ lookup-sum + 1x1 projections, depthwise 2x ConvTranspose upsample, 8-layer causal
// sliding-window transformer, SEANet decoder (modeling_mimi.py 5.5.0: _decode_frame :1613-1631,
// MimiSplitResidualVectorQuantizer.decode :1340-1350, MimiTransformerLayer :966-993,
// MimiAttention :681-738, MimiDecoder :1143-1173, MimiConvTranspose1d :402-409, MimiConv1d :331-351,
// MimiResnetBlock :437-451).
//
// Round-1 design: every dense block is an implicit GEMM over channel-last
// activations [T, C]: Linear, causal Conv1d (K = taps x Cin gathered from shifted rows) and causal
// ConvTranspose1d (stride s, kernel 2s == a 2-tap conv producing s*Cout columns, which IS the
// channel-last upsampled tensor), with ELU fused on the operand load and bias / GELU / LayerScale
// residual fused in the epilogue.  This first version runs the contractions in fp32 on the FFMA2
// pipe (parity 1e-4 against the fp32 oracle).  SOPRO_MIMI_BF16_TC mode (mimi_tc.cuh) runs every
// contraction whose channel count allows it on the tcgen05 tensor cores with bf16 operands, fp32
// accumulation in tensor memory and the same fused epilogues; the fp32 kernels stay for the exact mode
// and for the few narrow layers (Cin < 64).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sopro_b200.h"
#include "mimi_tc.cuh"

namespace mimi {

// shared with ar_engine.cu through sopro_last_error()
void set_error(const char* msg);

enum { EPI_NONE = 0, EPI_GELU = 1, EPI_RES_SCALE = 2, EPI_RES = 3 };

struct GemmOp {
  const float* A;   // channel-last input [B][Min][Cin]
  const float* W;   // [N][K], K = taps*Cin ordered (tap, ci)
  const float* bias;  // [bias_mod] or null
  const float* R;     // residual [B][M][N] or null
  const float* scale; // [N] LayerScale or null
  float* C;           // [B][M][ldc]
  long long a_bs, c_bs, r_bs;  // batch strides (floats)
  int M, N, K, Min, Cin, taps, dil, pad, ldc, bias_mod, epi, a_elu;
};

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// C[m][n] = epi( sum_k A'[m][k] * W[n][k] + bias ),  A'[m][(j,ci)] = act(X[m + j*dil - pad][ci]) (0 outside)
template <int BN>
__global__ void __launch_bounds__(256) igemm_kernel(const GemmOp op) {
  constexpr int BM = 64, BK = 16, TM = 4, TN = BN / 16;
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, b = blockIdx.z;
  const float* X = op.A + (size_t)b * op.a_bs;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // loader mapping: 64 rows x 4 float4 along k
  const int ty = tid >> 4, tx = tid & 15;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto load_a = [&](int k0) -> float4 {
    const int kk = k0 + lk;
    const int j = kk / op.Cin, ci = kk - j * op.Cin;
    const int m = m0 + lrow;
    const int rin = m + j * op.dil - op.pad;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < op.M && rin >= 0 && rin < op.Min) {
      v = *reinterpret_cast<const float4*>(X + (size_t)rin * op.Cin + ci);
      if (op.a_elu) {
        v.x = elu1(v.x);
        v.y = elu1(v.y);
        v.z = elu1(v.z);
        v.w = elu1(v.w);
      }
    }
    return v;
  };
  auto load_b = [&](int k0) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lrow < BN && n0 + lrow < op.N) v = __ldg(reinterpret_cast<const float4*>(op.W + (size_t)(n0 + lrow) * op.K + k0 + lk));
    return v;
  };
  auto store_tiles = [&](int buf, const float4& a, const float4& bq) {
    As[buf][lk + 0][lrow] = a.x;
    As[buf][lk + 1][lrow] = a.y;
    As[buf][lk + 2][lrow] = a.z;
    As[buf][lk + 3][lrow] = a.w;
    if (lrow < BN) {
      Bs[buf][lk + 0][lrow] = bq.x;
      Bs[buf][lk + 1][lrow] = bq.y;
      Bs[buf][lk + 2][lrow] = bq.z;
      Bs[buf][lk + 3][lrow] = bq.w;
    }
  };
  float4 ra = load_a(0), rb = load_b(0);
  store_tiles(0, ra, rb);
  __syncthreads();
  const int nk = op.K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) {
      ra = load_a((kt + 1) * BK);
      rb = load_b((kt + 1) * BK);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float bv[TN];
      if (TN == 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN]);
        bv[0] = b4.x;
        bv[1] = b4.y;
        bv[TN - 2] = b4.z;
        bv[TN - 1] = b4.w;
      } else {
        const float2 b2 = *reinterpret_cast<const float2*>(&Bs[buf][k][tx * TN]);
        bv[0] = b2.x;
        bv[TN - 1] = b2.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1, ra, rb);
    __syncthreads();
  }
  float* Cb = op.C + (size_t)b * op.c_bs;
  const float* Rb = op.R ? op.R + (size_t)b * op.r_bs : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= op.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= op.N) continue;
      float v = acc[i][j];
      if (op.bias) v += __ldg(op.bias + (n % op.bias_mod));
      if (op.epi == EPI_GELU) v = gelu_erf(v);
      else if (op.epi == EPI_RES_SCALE) v = Rb[(size_t)m * op.N + n] + __ldg(op.scale + n) * v;
      else if (op.epi == EPI_RES) v = Rb[(size_t)m * op.N + n] + v;
      Cb[(size_t)m * op.ldc + n] = v;
    }
  }
}

// RVQ lookup-sum: codes [B][Q][T] -> S [B][T][2*Dc] = [semantic sum | acoustic sum]
// A code outside [0, vocab) (an uncut EOS id, a negative pad) is clamped and recorded in *bad (sticky, read by
// sopro_mimi_check): the gather never leaves the table.  The reference's embedding lookup raises IndexError there.
__global__ void rvq_gather_kernel(const int* __restrict__ codes, const float* __restrict__ embed, float* __restrict__ S,
                                  int Q, int T, int Dc, int vocab, int n_sem, int* __restrict__ bad) {
  const int t = blockIdx.x, b = blockIdx.y;
  for (int c = threadIdx.x; c < Dc; c += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    for (int q = 0; q < Q; ++q) {
      int code = codes[((size_t)b * Q + q) * T + t];
      if (code < 0 || code >= vocab) {
        if (c == 0) atomicOr(bad, 1);
        code = min(max(code, 0), vocab - 1);
      }
      const float e = __ldg(embed + ((size_t)q * vocab + code) * Dc + c);
      if (q < n_sem) s0 += e;
      else s1 += e;
    }
    float* o = S + ((size_t)b * T + t) * (2 * Dc);
    o[c] = s0;
    o[Dc + c] = s1;
  }
}

// depthwise ConvTranspose k=4 s=2, causal: y[2t+r][c] = x[t][c]*w[c][r] + x[t-1][c]*w[c][r+2]
// `prev` (streaming, B = 1): the frame before x[0] (zeros at the start of a stream), else null
__global__ void upsample_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int T,
                                int C, const float* __restrict__ prev) {
  const int to = blockIdx.x, b = blockIdx.y;  // output row 0..2T-1
  const int t = to >> 1, r = to & 1;
  const float* xb = x + (size_t)b * T * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = xb[(size_t)t * C + c] * __ldg(w + c * 4 + r);
    if (t > 0) v += xb[(size_t)(t - 1) * C + c] * __ldg(w + c * 4 + r + 2);
    else if (prev) v += prev[c] * __ldg(w + c * 4 + r + 2);
    y[((size_t)b * 2 * T + to) * C + c] = v;
  }
}

// LayerNorm over C (one warp per row)
__device__ __forceinline__ void put(float* p, float v) { *p = v; }
__device__ __forceinline__ void put(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename OutT>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bb,
                                 OutT* __restrict__ y, long long rows, int C, float eps) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mean;
    v += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float inv = 1.0f / sqrtf(v / (float)C + eps);
  for (int c = lane; c < C; c += 32) put(y + row * C + c, (xr[c] - mean) * inv * __ldg(w + c) + __ldg(bb + c));
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  reinterpret_cast<uint2*>(y)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// RoPE in place on the q and k thirds of QKV [rows][3C]; rope table [T2][Dh/2] cos, then sin
// Streaming (kring != null, B = 1): row t sits at absolute position pos0 + t; its rotated key and its value are also
// appended to the layer's K/V ring at slot (pos0 + t) % R.
__global__ void rope_kernel(float* __restrict__ qkv, const float* __restrict__ cs, int T2, int tab_T2, int C, int H,
                            int pos0, float* __restrict__ kring, float* __restrict__ vring, int R) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int Dh = C / H, half = Dh / 2;
  float* row = qkv + ((size_t)b * T2 + t) * 3 * C;
  const float* cosr = cs + (size_t)(pos0 + t) * half;
  const float* sinr = cs + (size_t)(tab_T2 + pos0 + t) * half;  // table: [cos rows 0..tab_T2) | sin rows 0..tab_T2)]
  for (int i = threadIdx.x; i < 2 * H * half; i += blockDim.x) {
    const int which = i / (H * half);  // 0 = q, 1 = k
    const int rem = i - which * H * half;
    const int h = rem / half, d = rem - h * half;
    float* p = row + which * C + h * Dh;
    const float x1 = p[d], x2 = p[d + half];
    const float c = cosr[d], s = sinr[d];
    p[d] = x1 * c - x2 * s;          // q*cos + rotate_half(q)*sin, first half: -x2
    p[d + half] = x2 * c + x1 * s;   // second half: +x1
  }
  if (kring) {
    __syncthreads();
    const size_t slot = (size_t)((pos0 + t) % R) * C;
    for (int c4 = threadIdx.x * 4; c4 < C; c4 += blockDim.x * 4) {
      *reinterpret_cast<float4*>(kring + slot + c4) = *reinterpret_cast<const float4*>(row + C + c4);
      *reinterpret_cast<float4*>(vring + slot + c4) = *reinterpret_cast<const float4*>(row + 2 * C + c4);
    }
  }
}

// Tensor-core attention operands from the fp32 QKV rows [B][T2][3C]: rotated q and k as bf16 [B][T2][C], v transposed
// as bf16 [B][C][T2p] (keys contiguous: the K-major B operand of P.V).  32 rows per block.
template <int DH>
__global__ void __launch_bounds__(256) rope_pack_kernel(const float* __restrict__ qkv, const float* __restrict__ cs,
                                                        __nv_bfloat16* __restrict__ qh, __nv_bfloat16* __restrict__ kh,
                                                        __nv_bfloat16* __restrict__ vt, int T2, long long T2p, int tab_T2, int C, int H) {
  extern __shared__ __nv_bfloat16 sv[];  // [32][C + 2]
  constexpr int half = DH / 2;
  const int t0 = blockIdx.x * 32, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = min(32, T2 - t0);
  for (int tt = warp; tt < rows; tt += 8) {  // one warp per row
    const int t = t0 + tt;
    const float* row = qkv + ((size_t)b * T2 + t) * 3 * C;
    const float* cosr = cs + (size_t)t * half;
    const float* sinr = cs + (size_t)(tab_T2 + t) * half;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* o = (which ? kh : qh) + ((size_t)b * T2 + t) * C;
      for (int i = lane; i < H * half; i += 32) {
        const int h = i / half, d = i % half;
        const float x1 = row[which * C + h * DH + d], x2 = row[which * C + h * DH + d + half];
        const float c = cosr[d], s = sinr[d];
        o[h * DH + d] = __float2bfloat16_rn(x1 * c - x2 * s);
        o[h * DH + d + half] = __float2bfloat16_rn(x2 * c + x1 * s);
      }
    }
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(row + 2 * C + c);
      __nv_bfloat16* d = sv + tt * (C + 2) + c;
      d[0] = __float2bfloat16_rn(v.x);
      d[1] = __float2bfloat16_rn(v.y);
      d[2] = __float2bfloat16_rn(v.z);
      d[3] = __float2bfloat16_rn(v.w);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * C; idx += blockDim.x) {
    const int c = idx >> 5, tt = idx & 31;
    if (t0 + tt < T2p) vt[((size_t)b * C + c) * T2p + t0 + tt] = tt < rows ? sv[tt * (C + 2) + c] : __float2bfloat16_rn(0.f);  // pad stays finite
  }
}

// causal sliding-window attention, one warp per (b, h, query); QKV rotated; out [B][T2][C]
// Streaming (kring != null, B = 1): query row i sits at absolute position pos0 + i and the keys / values of positions
// [pos - window + 1, pos] are read from the layer's ring (slot = position % R); the arithmetic and its order are the
// full decode's, so a chunked decode equals the full decode's prefix bit for bit in fp32 mode.
template <typename OutT>
__global__ void __launch_bounds__(256) attn_kernel(const float* __restrict__ qkv, OutT* __restrict__ out, int T2, int C,
                                                   int H, int window, int pos0, const float* __restrict__ kring,
                                                   const float* __restrict__ vring, int R) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Dh = C / H;
  const int i = blockIdx.x * 8 + warp, h = blockIdx.y, b = blockIdx.z;
  float* qs = sm + warp * (Dh + window);
  float* sc = qs + Dh;
  if (i >= T2) return;
  const float* base = qkv + (size_t)b * T2 * 3 * C;
  const float* q = base + (size_t)i * 3 * C + h * Dh;
  for (int d = lane; d < Dh; d += 32) qs[d] = q[d];
  __syncwarp();
  const int ia = pos0 + i;  // absolute position
  const int j0 = max(0, ia - window + 1);
  const int nk = ia - j0 + 1;
  const float scale = 1.0f / sqrtf((float)Dh);
  float mx = -INFINITY;
  for (int jj = lane; jj < nk; jj += 32) {
    const float* kr = kring ? kring + (size_t)((j0 + jj) % R) * C + h * Dh : base + (size_t)(j0 + jj) * 3 * C + C + h * Dh;
    float s = 0.f;
    for (int d = 0; d < Dh; d += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(kr + d);
      s += kk.x * qs[d] + kk.y * qs[d + 1] + kk.z * qs[d + 2] + kk.w * qs[d + 3];
    }
    s *= scale;
    sc[jj] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int jj = lane; jj < nk; jj += 32) {
    const float e = expf(sc[jj] - mx);
    sc[jj] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncwarp();
  const float inv = 1.0f / sum;
  for (int d = lane; d < Dh; d += 32) {
    float o = 0.f;
    if (vring) {
      for (int jj = 0; jj < nk; ++jj) o += (sc[jj] * inv) * vring[(size_t)((j0 + jj) % R) * C + h * Dh + d];
    } else {
      for (int jj = 0; jj < nk; ++jj) o += (sc[jj] * inv) * base[(size_t)(j0 + jj) * 3 * C + 2 * C + h * Dh + d];
    }
    put(out + ((size_t)b * T2 + i) * C + h * Dh + d, o);
  }
}

// final conv: ELU -> causal conv k taps, Cin -> 1
// rows r >= lo are readable (lo = 0: the causal zero pad; streaming: lo = -(taps-1), the carried context rows sit in
// front of x)
__global__ void final_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ y, long long Tn, int Cin, int taps, int lo) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= Tn) return;
  const float* xb = x + (size_t)b * Tn * Cin;
  float acc = __ldg(bias);
  for (int j = 0; j < taps; ++j) {
    const long long r = t + j - (taps - 1);
    if (r < lo) continue;
    const float* xr = xb + r * Cin;
    for (int c = 0; c < Cin; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + j * Cin + c));
      acc += elu1(v.x) * ww.x + elu1(v.y) * ww.y + elu1(v.z) * ww.z + elu1(v.w) * ww.w;
    }
  }
  y[(size_t)b * Tn + t] = acc;
}

// final conv on the bf16 activation [B][Tn][Cin] that already went through ELU (tensor-core mode): one thread per
// input row computes the row's dot product with every tap's weights (each row is read once), the taps are combined
// through shared memory.  256 - (taps-1) outputs per block.
__global__ void __launch_bounds__(256) final_conv_h_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ y, long long Tn,
                                                           int Cin, int taps, int lo) {
  extern __shared__ float fsm[];  // [taps][256] partial dots, then [taps*Cin] weights
  float* sp = fsm;
  float* sw = fsm + taps * 256;
  const int tid = threadIdx.x, halo = taps - 1, per = 256 - halo, b = blockIdx.y;
  for (int i = tid; i < taps * Cin; i += 256) sw[i] = __ldg(w + i);
  __syncthreads();
  const long long r = (long long)blockIdx.x * per - halo + tid;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (r >= lo && r < Tn) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + ((long long)b * Tn + r) * Cin);
    for (int c = 0; c < Cin; c += 8) {
      const uint4 v = xr[c >> 3];
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
      float f[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        f[2 * e] = __uint_as_float(u[e] << 16);
        f[2 * e + 1] = __uint_as_float(u[e] & 0xffff0000u);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < taps) {
          const float* wj = sw + j * Cin + c;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[j] = fmaf(f[e], wj[e], acc[j]);
        }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < taps) sp[j * 256 + tid] = acc[j];
  __syncthreads();
  if (tid >= halo && r < Tn) {  // output sample r: tap j reads row r + j - halo
    float o = __ldg(bias);
    for (int j = 0; j < taps; ++j) o += sp[j * 256 + tid - halo + j];
    y[(size_t)b * Tn + r] = o;
  }
}

}  // namespace mimi

using namespace mimi;

namespace {
int mfail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  mimi::set_error(buf);
  return code;
}
#define MCK(call)                                                                                      \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return mfail(SOPRO_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

uint16_t bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN stays NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// bf16 copies of the GEMM weight matrices (tensor-core mode); offsets in elements, 256-B aligned
struct Bf16Arena {
  std::vector<uint16_t> host;
  size_t add(const float* p, size_t n) {
    const size_t off = (host.size() + 127) / 128 * 128;
    host.resize(off + n);
    for (size_t i = 0; i < n; ++i) host[off + i] = bf16_rne(p[i]);
    return off;
  }
};

struct DevArena {
  std::vector<float> host;
  size_t add(const float* p, size_t n) {
    const size_t off = (host.size() + 63) / 64 * 64;
    host.resize(off + n);
    if (p) memcpy(host.data() + off, p, n * 4);
    return off;
  }
};
}  // namespace

struct sopro_mimi {
  int device = 0;
  sopro_mimi_config_t cfg{};
  float* dev = nullptr;
  size_t n_floats = 0;
  // offsets (floats) into dev
  size_t embed = 0, rvq_w = 0, up_w = 0;
  struct Layer {
    size_t ln1w, ln1b, qkv, wo, ls1, ln2w, ln2b, fc1, fc2, ls2;
    size_t qkv_h, wo_h, fc1_h, fc2_h;  // bf16 arena
  };
  std::vector<Layer> layers;
  size_t c0w = 0, c0b = 0, c0w_h = 0;
  struct Stage {
    size_t tw, tb, r1w, r1b, r2w, r2b;
    size_t tw_h, r1w_h, r2w_h;  // bf16 arena
    int ratio, cin, cout;
  };
  __nv_bfloat16* dev_h = nullptr;  // bf16 weight arena
  int precision = SOPRO_MIMI_BF16_TC;
  std::vector<Stage> stages;
  size_t lw = 0, lb = 0;
  // rope table + workspace
  float* rope = nullptr;
  int rope_T2 = 0;
  float* ws = nullptr;
  size_t ws_bytes = 0;
  int* codes_dev = nullptr;
  size_t codes_cap = 0;
  int* bad_code = nullptr;  // sticky flag: a decode saw a code outside [0, vocab)
  // launch-bound small decodes (streaming chunks, time-to-first-audio) are replayed from CUDA graphs captured over
  // internal static buffers; every graph dies when the workspace or the rope table is reallocated
  struct Replay {
    int B, T, precision;
    cudaGraphExec_t exec;
  };
  std::vector<Replay> replays;
  int* g_codes = nullptr;   // [kGraphFrames * n_q]
  float* g_wav = nullptr;   // [kGraphFrames * hop]
  bool graphs = true;
  cudaStream_t cap_stream = nullptr;
};

extern "C" {

int sopro_mimi_create(const sopro_mimi_config_t* cfg, const sopro_mimi_weights_t* w, int device, sopro_mimi_t** out) {
  if (!cfg || !w || !out) return mfail(SOPRO_ERR_INVALID, "null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0) return mfail(SOPRO_ERR_UNSUPPORTED, "no CUDA device; the Mimi decoder has no CPU fallback");
  if (device < 0 || device >= ndev) return mfail(SOPRO_ERR_INVALID, "device %d out of range", device);
  cudaDeviceProp prop;
  MCK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return mfail(SOPRO_ERR_UNSUPPORTED, "device is sm_%d%d; this build targets sm_100a only", prop.major, prop.minor);
  const int C = cfg->hidden, Dc = cfg->codebook_dim, Q = cfg->n_q, V = cfg->vocab, NL = cfg->n_layers, FF = cfg->ffn;
  if (C % 64 || Dc % 4 || C != 2 * Dc || NL < 1 || cfg->n_ratios < 1 || cfg->n_ratios > 8 || cfg->n_heads < 1 || C % cfg->n_heads ||
      (C / cfg->n_heads) % 4 || FF % 16)
    return mfail(SOPRO_ERR_INVALID, "unsupported Mimi geometry (hidden=%d codebook_dim=%d)", C, Dc);
  MCK(cudaSetDevice(device));
  sopro_mimi* m = new sopro_mimi();
  m->device = device;
  m->cfg = *cfg;
  DevArena A;
  Bf16Arena Hh;
  m->embed = A.add(w->embed, (size_t)Q * V * Dc);
  {  // [C][2*Dc] = [W_sem | W_ac]
    std::vector<float> cat((size_t)C * 2 * Dc);
    for (int n = 0; n < C; ++n)
      for (int k = 0; k < Dc; ++k) {
        cat[(size_t)n * 2 * Dc + k] = w->sem_out_proj[(size_t)n * Dc + k];
        cat[(size_t)n * 2 * Dc + Dc + k] = w->ac_out_proj[(size_t)n * Dc + k];
      }
    m->rvq_w = A.add(cat.data(), cat.size());
  }
  m->up_w = A.add(w->upsample_w, (size_t)C * 4);
  for (int l = 0; l < NL; ++l) {
    const sopro_mimi_layer_weights_t& L = w->layer[l];
    sopro_mimi::Layer d;
    d.ln1w = A.add(L.ln1_w, C);
    d.ln1b = A.add(L.ln1_b, C);
    std::vector<float> qkv((size_t)3 * C * C);
    memcpy(qkv.data(), L.q_w, (size_t)C * C * 4);
    memcpy(qkv.data() + (size_t)C * C, L.k_w, (size_t)C * C * 4);
    memcpy(qkv.data() + (size_t)2 * C * C, L.v_w, (size_t)C * C * 4);
    d.qkv = A.add(qkv.data(), qkv.size());
    d.qkv_h = Hh.add(qkv.data(), qkv.size());
    d.wo = A.add(L.o_w, (size_t)C * C);
    d.wo_h = Hh.add(L.o_w, (size_t)C * C);
    d.fc1_h = Hh.add(L.fc1_w, (size_t)FF * C);
    d.fc2_h = Hh.add(L.fc2_w, (size_t)C * FF);
    d.ls1 = A.add(L.ls1, C);
    d.ln2w = A.add(L.ln2_w, C);
    d.ln2b = A.add(L.ln2_b, C);
    d.fc1 = A.add(L.fc1_w, (size_t)FF * C);
    d.fc2 = A.add(L.fc2_w, (size_t)C * FF);
    d.ls2 = A.add(L.ls2, C);
    m->layers.push_back(d);
  }
  // conv weights [Cout][Cin][k] -> [Cout][(tap, ci)]
  auto repack_conv = [&](const float* src, int cout, int cin, int k, size_t* half_off) {
    std::vector<float> r((size_t)cout * k * cin);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int j = 0; j < k; ++j) r[((size_t)co * k + j) * cin + ci] = src[((size_t)co * cin + ci) * k + j];
    if (half_off) *half_off = Hh.add(r.data(), r.size());
    return A.add(r.data(), r.size());
  };
  int ch = cfg->num_filters << cfg->n_ratios;  // 1024
  m->c0w = repack_conv(w->conv0_w, ch, C, cfg->kernel, &m->c0w_h);
  m->c0b = A.add(w->conv0_b, ch);
  for (int s = 0; s < cfg->n_ratios; ++s) {
    const sopro_mimi_stage_weights_t& S = w->stage[s];
    sopro_mimi::Stage d;
    const int r = cfg->ratios[s], cin = ch, cout = ch / 2;
    d.ratio = r;
    d.cin = cin;
    d.cout = cout;
    // ConvTranspose weight [Cin][Cout][2r] -> [(phase, co)][(tap, ci)]: tap 0 <-> x[t-1] <-> w[.., phase + r], tap 1 <-> x[t] <-> w[.., phase]
    std::vector<float> tw((size_t)r * cout * 2 * cin);
    for (int ph = 0; ph < r; ++ph)
      for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci) {
          const size_t n = (size_t)ph * cout + co;
          tw[(n * 2 + 0) * cin + ci] = S.convt_w[((size_t)ci * cout + co) * 2 * r + ph + r];
          tw[(n * 2 + 1) * cin + ci] = S.convt_w[((size_t)ci * cout + co) * 2 * r + ph];
        }
    d.tw = A.add(tw.data(), tw.size());
    d.tw_h = Hh.add(tw.data(), tw.size());
    d.tb = A.add(S.convt_b, cout);
    d.r1w = repack_conv(S.res1_w, cout / cfg->compress, cout, cfg->res_kernel, &d.r1w_h);
    d.r1b = A.add(S.res1_b, cout / cfg->compress);
    d.r2w = repack_conv(S.res2_w, cout, cout / cfg->compress, 1, &d.r2w_h);
    d.r2b = A.add(S.res2_b, cout);
    m->stages.push_back(d);
    ch = cout;
  }
  m->lw = repack_conv(w->last_w, 1, ch, cfg->last_kernel, nullptr);
  m->lb = A.add(w->last_b, 1);
  m->n_floats = A.host.size();
  cudaError_t err = cudaMalloc(&m->dev, m->n_floats * 4);
  if (err == cudaSuccess) err = cudaMemcpy(m->dev, A.host.data(), m->n_floats * 4, cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMalloc(&m->bad_code, 256);
  if (err == cudaSuccess) err = cudaMemset(m->bad_code, 0, 256);
  if (err == cudaSuccess) err = cudaMalloc(&m->dev_h, Hh.host.size() * 2);
  if (err == cudaSuccess) err = cudaMemcpy(m->dev_h, Hh.host.data(), Hh.host.size() * 2, cudaMemcpyHostToDevice);
  if (err == cudaSuccess && !tc::encode_tiled_fn()) {
    cudaFree(m->dev);
    cudaFree(m->dev_h);
    delete m;
    return mfail(SOPRO_ERR_UNSUPPORTED, "driver has no cuTensorMapEncodeTiled entry point");
  }
  if (err != cudaSuccess) {
    if (m->dev) cudaFree(m->dev);
    if (m->dev_h) cudaFree(m->dev_h);
    delete m;
    return mfail(SOPRO_ERR_CUDA, "Mimi weight upload failed: %s", cudaGetErrorString(err));
  }
  *out = m;
  return SOPRO_OK;
}

int sopro_mimi_destroy(sopro_mimi_t* m) {
  if (!m) return SOPRO_OK;
  cudaSetDevice(m->device);
  cudaFree(m->dev);
  cudaFree(m->dev_h);
  cudaFree(m->rope);
  cudaFree(m->ws);
  cudaFree(m->codes_dev);
  cudaFree(m->bad_code);
  for (auto& r : m->replays) cudaGraphExecDestroy(r.exec);
  cudaFree(m->g_codes);
  cudaFree(m->g_wav);
  if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
  delete m;
  return SOPRO_OK;
}

int64_t sopro_mimi_samples_per_frame(const sopro_mimi_t* m) {
  if (!m) return 0;
  int64_t s = 2;
  for (int i = 0; i < m->cfg.n_ratios; ++i) s *= m->cfg.ratios[i];
  return s;
}

static int launch_gemm(const GemmOp& op, int B, cudaStream_t st) {
  if (op.K % 16 || op.Cin % 4) return mfail(SOPRO_ERR_INVALID, "igemm: K=%d Cin=%d not aligned", op.K, op.Cin);
  if (op.N % 64 == 0 || op.N > 32) {
    dim3 grid((op.M + 63) / 64, (op.N + 63) / 64, B);
    igemm_kernel<64><<<grid, 256, 0, st>>>(op);
  } else {
    dim3 grid((op.M + 63) / 64, (op.N + 31) / 32, B);
    igemm_kernel<32><<<grid, 256, 0, st>>>(op);
  }
  MCK(cudaGetLastError());
  return SOPRO_OK;
}

int sopro_mimi_set_precision(sopro_mimi_t* m, int precision) {
  if (!m) return mfail(SOPRO_ERR_INVALID, "null argument");
  if (precision != SOPRO_MIMI_FP32 && precision != SOPRO_MIMI_BF16_TC) return mfail(SOPRO_ERR_INVALID, "unknown precision %d", precision);
  m->precision = precision;
  return SOPRO_OK;
}

}  // extern "C"

namespace {
constexpr long long kGraphFrames = 64;  // decodes of at most this many frames (B*T) go through the graph cache

void drop_replays(sopro_mimi* m) {
  for (auto& r : m->replays) cudaGraphExecDestroy(r.exec);
  m->replays.clear();
}

// Allocations and table uploads a decode of [B, T] needs; never called inside a stream capture.
// RoPE table [cos rows 0..T2) | sin rows 0..T2)] covering at least T2 positions
int ensure_rope(sopro_mimi* m, int T2, cudaStream_t st) {
  const sopro_mimi_config_t& c = m->cfg;
  const int Dh = c.hidden / c.n_heads;
  if (m->rope_T2 < T2) {
    drop_replays(m);
    cudaFree(m->rope);
    m->rope = nullptr;
    m->rope_T2 = 0;
    std::vector<float> tab((size_t)2 * T2 * (Dh / 2));
    for (int t = 0; t < T2; ++t)
      for (int d = 0; d < Dh / 2; ++d) {
        const float inv = 1.0f / powf(c.rope_theta, (float)(2 * d) / (float)Dh);
        const float f = (float)t * inv;
        tab[(size_t)t * (Dh / 2) + d] = cosf(f);
        tab[(size_t)(T2 + t) * (Dh / 2) + d] = sinf(f);
      }
    MCK(cudaMalloc(&m->rope, tab.size() * 4));
    MCK(cudaMemcpyAsync(m->rope, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, st));
    MCK(cudaStreamSynchronize(st));
    m->rope_T2 = T2;
  }
  return SOPRO_OK;
}

int mimi_prepare(sopro_mimi* m, int B, int T, cudaStream_t st) {
  const sopro_mimi_config_t& c = m->cfg;
  const int C = c.hidden, T2 = 2 * T, FF = c.ffn;
  {
    // streams share the table: never shrink it, grow with headroom
    const int rc = ensure_rope(m, T2, st);
    if (rc) return rc;
  }
  long long up = 2;
  for (int i = 0; i < c.n_ratios; ++i) up *= c.ratios[i];
  const size_t big = (size_t)B * T * up * c.num_filters;
  const size_t tr = (size_t)B * T2 * (size_t)std::max(3 * C, FF);
  const size_t bufsz = (std::max(std::max(big, tr), (size_t)B * T2 * (c.num_filters << c.n_ratios)) + 63) / 64 * 64;
  const size_t xsz = ((size_t)B * T2 * C + 63) / 64 * 64;
  const size_t need = (3 * bufsz + 2 * xsz) * 4 + 3 * bufsz * 2;
  if (m->ws_bytes < need) {
    drop_replays(m);
    cudaFree(m->ws);
    m->ws = nullptr;
    m->ws_bytes = 0;
    cudaError_t e = cudaMalloc(&m->ws, need);
    if (e != cudaSuccess) return mfail(SOPRO_ERR_CUDA, "Mimi workspace %zu MB: %s", need >> 20, cudaGetErrorString(e));
    m->ws_bytes = need;
  }
  return SOPRO_OK;
}

int mimi_enqueue(sopro_mimi* m, const int32_t* codes, int B, int T, float* wav, cudaStream_t st);
}  // namespace

extern "C" {

int sopro_mimi_decode(sopro_mimi_t* m, const int32_t* codes, int B, int T, float* wav, void* stream) {
  if (!m || !codes || !wav) return mfail(SOPRO_ERR_INVALID, "null argument");
  if (B < 1 || T < 1) return mfail(SOPRO_ERR_INVALID, "B and T must be >= 1");
  if (B > 65535) return mfail(SOPRO_ERR_INVALID, "B must be <= 65535");
  MCK(cudaSetDevice(m->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = mimi_prepare(m, B, T, st);
  if (rc) return rc;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  MCK(cudaStreamIsCapturing(st, &cap));
  if (!m->graphs || (long long)B * T > kGraphFrames || cap != cudaStreamCaptureStatusNone) return mimi_enqueue(m, codes, B, T, wav, st);
  const size_t nc = (size_t)B * m->cfg.n_q * T, nw = (size_t)B * T * (size_t)sopro_mimi_samples_per_frame(m);
  if (!m->g_codes) {
    MCK(cudaMalloc(&m->g_codes, (size_t)kGraphFrames * m->cfg.n_q * 4));
    MCK(cudaMalloc(&m->g_wav, (size_t)kGraphFrames * (size_t)sopro_mimi_samples_per_frame(m) * 4));
  }
  cudaGraphExec_t exec = nullptr;
  for (auto& r : m->replays)
    if (r.B == B && r.T == T && r.precision == m->precision) exec = r.exec;
  if (!exec) {
    if (m->replays.size() >= 32) drop_replays(m);
    // captured on a private stream (the caller's may be the legacy default stream, which cannot capture)
    if (!m->cap_stream) MCK(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
    MCK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
    rc = mimi_enqueue(m, m->g_codes, B, T, m->g_wav, m->cap_stream);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(m->cap_stream, &graph);
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (ce != cudaSuccess) return mfail(SOPRO_ERR_CUDA, "Mimi graph capture: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return mfail(SOPRO_ERR_CUDA, "Mimi graph instantiate: %s", cudaGetErrorString(ce));
    m->replays.push_back({B, T, m->precision, exec});
  }
  MCK(cudaMemcpyAsync(m->g_codes, codes, nc * 4, cudaMemcpyDeviceToDevice, st));
  MCK(cudaGraphLaunch(exec, st));
  MCK(cudaMemcpyAsync(wav, m->g_wav, nw * 4, cudaMemcpyDeviceToDevice, st));
  return SOPRO_OK;
}

}  // extern "C"

namespace {
int mimi_enqueue(sopro_mimi* m, const int32_t* codes, int B, int T, float* wav, cudaStream_t st) {
  const sopro_mimi_config_t& c = m->cfg;
  const int C = c.hidden, T2 = 2 * T, H = c.n_heads, Dh = C / H, FF = c.ffn;
  const float* Wd = m->dev;
  const __nv_bfloat16* Wh = m->dev_h;
  const bool use_tc = m->precision == SOPRO_MIMI_BF16_TC;
  // ---- workspace (mimi_prepare): three fp32 ping-pong buffers sized for the widest SEANet activation + transformer
  //      scratch, the residual stream and its normalised copy, three bf16 buffers for the tensor-core operands
  long long up = 2;
  for (int i = 0; i < c.n_ratios; ++i) up *= c.ratios[i];
  const size_t big = (size_t)B * T * up * c.num_filters;                     // [T*1920][64]
  const size_t tr = (size_t)B * T2 * (size_t)std::max(3 * C, FF);            // QKV / MLP hidden
  const size_t bufsz = (std::max(std::max(big, tr), (size_t)B * T2 * (c.num_filters << c.n_ratios)) + 63) / 64 * 64;
  const size_t xsz = ((size_t)B * T2 * C + 63) / 64 * 64;
  float* b0 = m->ws;
  float* b1 = b0 + bufsz;
  float* b2 = b1 + bufsz;
  float* x = b2 + bufsz;    // residual stream [B][T2][C]
  float* ln = x + xsz;      // normalised copy (fp32 mode)
  __nv_bfloat16* h0 = reinterpret_cast<__nv_bfloat16*>(ln + xsz);
  __nv_bfloat16* h1 = h0 + bufsz;
  __nv_bfloat16* h2 = h1 + bufsz;
  // ---- RVQ + projection + upsample (small; fp32 in both modes)
  rvq_gather_kernel<<<dim3(T, B), 256, 0, st>>>(codes, Wd + m->embed, b0, c.n_q, T, c.codebook_dim, c.vocab, c.n_sem, m->bad_code);
  MCK(cudaGetLastError());
  GemmOp g{};
  auto lin = [&](const float* A, int M, int K, const float* W, int N, float* Cc, int epi, const float* R, const float* scale) {
    g = GemmOp{};
    g.A = A; g.W = W; g.C = Cc; g.R = R; g.scale = scale; g.bias = nullptr;
    g.M = M; g.N = N; g.K = K; g.Min = M; g.Cin = K; g.taps = 1; g.dil = 1; g.pad = 0; g.ldc = N; g.bias_mod = N; g.epi = epi;
    g.a_bs = (long long)M * K; g.c_bs = (long long)M * N; g.r_bs = (long long)M * N;
    return launch_gemm(g, B, st);
  };
  auto conv = [&](const float* A, long long Tin, int cin, int taps, int pad, const float* W, const float* bias, int N, int bias_mod,
                  float* Cc, int elu, int epi, const float* R) {
    g = GemmOp{};
    g.A = A; g.W = W; g.C = Cc; g.R = R; g.bias = bias; g.scale = nullptr;
    g.M = (int)Tin; g.N = N; g.K = taps * cin; g.Min = (int)Tin; g.Cin = cin; g.taps = taps; g.dil = 1; g.pad = pad; g.ldc = N;
    g.bias_mod = bias_mod; g.epi = epi; g.a_elu = elu;
    g.a_bs = Tin * cin; g.c_bs = Tin * N; g.r_bs = Tin * N;
    return launch_gemm(g, B, st);
  };
  // tensor-core implicit GEMM: X bf16 [B][rows][cin] (ELU already applied by its producer where the layer wants it)
  auto tcg = [&](const __nv_bfloat16* X, long long rows, int cin, int taps, int pad, const __nv_bfloat16* W, const float* bias,
                 int N, int bias_mod, int epi, const float* R, const float* scale, float* of, __nv_bfloat16* oh, int out_elu) {
    tc::TcOp o{};
    o.bias = bias; o.R = R; o.scale = scale; o.out_f32 = of; o.out_bf16 = oh;
    o.c_bs = rows * N; o.M = (int)rows; o.N = N; o.K = taps * cin; o.Cin = cin; o.dil = 1; o.pad = pad;
    o.bias_mod = bias_mod; o.epi = epi; o.out_elu = out_elu;
    cudaError_t e = tc::launch(X, rows, W, o, B, st);
    if (e != cudaSuccess) return mfail(SOPRO_ERR_CUDA, "tensor-core GEMM (N=%d K=%d): %s", N, o.K, cudaGetErrorString(e));
    return (int)SOPRO_OK;
  };
  int rc;
  if ((rc = lin(b0, T, C, Wd + m->rvq_w, C, b1, EPI_NONE, nullptr, nullptr))) return rc;
  upsample_kernel<<<dim3(T2, B), 256, 0, st>>>(b1, Wd + m->up_w, x, T, C, nullptr);
  MCK(cudaGetLastError());
  // ---- transformer
  const long long rows = (long long)B * T2;
  const unsigned ln_grid = (unsigned)((rows + 7) / 8);
  const size_t asm_bytes = (size_t)8 * (Dh + c.window) * 4;
  const bool tc_tr = use_tc && tc::supported(3 * C, C, C) && tc::supported(C, C, C) && tc::supported(FF, C, C) && tc::supported(C, FF, FF);
  const bool tc_attn = tc_tr && tc::attn_supported(C, H, c.window) && (size_t)32 * (C + 2) * 2 <= 48 * 1024;
  const long long T2p = (T2 + 7) / 8 * 8;  // v^T row pitch: tensor-map strides are multiples of 16 bytes
  for (const sopro_mimi::Layer& L : m->layers) {
    if (tc_tr) {
      layernorm_kernel<<<ln_grid, 256, 0, st>>>(x, Wd + L.ln1w, Wd + L.ln1b, h0, rows, C, c.norm_eps);
      if ((rc = tcg(h0, T2, C, 1, 0, Wh + L.qkv_h, nullptr, 3 * C, 3 * C, tc::EPI_NONE, nullptr, nullptr, b0, nullptr, 0))) return rc;
      __nv_bfloat16* att = reinterpret_cast<__nv_bfloat16*>(b1);  // b1 is idle during the transformer
      if (tc_attn) {
        // q -> h0 (the LayerNorm copy is dead), k -> h1, v^T -> h2
        rope_pack_kernel<tc::kAttnDh><<<dim3((T2 + 31) / 32, B), 256, (size_t)32 * (C + 2) * 2, st>>>(b0, m->rope, h0, h1, h2, T2, T2p, m->rope_T2, C, H);
        MCK(cudaGetLastError());
        cudaError_t ae = tc::launch_attn(h0, h1, h2, att, B, T2, T2p, C, H, c.window, st);
        if (ae != cudaSuccess) return mfail(SOPRO_ERR_CUDA, "tensor-core attention: %s", cudaGetErrorString(ae));
      } else {
        rope_kernel<<<dim3(T2, B), 256, 0, st>>>(b0, m->rope, T2, m->rope_T2, C, H, 0, nullptr, nullptr, 1);
        attn_kernel<<<dim3((T2 + 7) / 8, H, B), 256, asm_bytes, st>>>(b0, att, T2, C, H, c.window, 0, nullptr, nullptr, 1);
        MCK(cudaGetLastError());
      }
      if ((rc = tcg(att, T2, C, 1, 0, Wh + L.wo_h, nullptr, C, C, tc::EPI_RES_SCALE, x, Wd + L.ls1, x, nullptr, 0))) return rc;
      layernorm_kernel<<<ln_grid, 256, 0, st>>>(x, Wd + L.ln2w, Wd + L.ln2b, h0, rows, C, c.norm_eps);
      if ((rc = tcg(h0, T2, C, 1, 0, Wh + L.fc1_h, nullptr, FF, FF, tc::EPI_GELU, nullptr, nullptr, nullptr, h2, 0))) return rc;
      if ((rc = tcg(h2, T2, FF, 1, 0, Wh + L.fc2_h, nullptr, C, C, tc::EPI_RES_SCALE, x, Wd + L.ls2, x, nullptr, 0))) return rc;
    } else {
      layernorm_kernel<<<ln_grid, 256, 0, st>>>(x, Wd + L.ln1w, Wd + L.ln1b, ln, rows, C, c.norm_eps);
      if ((rc = lin(ln, T2, C, Wd + L.qkv, 3 * C, b0, EPI_NONE, nullptr, nullptr))) return rc;
      rope_kernel<<<dim3(T2, B), 256, 0, st>>>(b0, m->rope, T2, m->rope_T2, C, H, 0, nullptr, nullptr, 1);
      attn_kernel<<<dim3((T2 + 7) / 8, H, B), 256, asm_bytes, st>>>(b0, b1, T2, C, H, c.window, 0, nullptr, nullptr, 1);
      MCK(cudaGetLastError());
      if ((rc = lin(b1, T2, C, Wd + L.wo, C, x, EPI_RES_SCALE, x, Wd + L.ls1))) return rc;
      layernorm_kernel<<<ln_grid, 256, 0, st>>>(x, Wd + L.ln2w, Wd + L.ln2b, ln, rows, C, c.norm_eps);
      if ((rc = lin(ln, T2, C, Wd + L.fc1, FF, b0, EPI_GELU, nullptr, nullptr))) return rc;
      if ((rc = lin(b0, T2, FF, Wd + L.fc2, C, x, EPI_RES_SCALE, x, Wd + L.ls2))) return rc;
    }
  }
  // ---- SEANet decoder
  long long Tn = T2;
  int ch = c.num_filters << c.n_ratios;
  if (!use_tc) {
    if ((rc = conv(x, Tn, C, c.kernel, c.kernel - 1, Wd + m->c0w, Wd + m->c0b, ch, ch, b0, 0, EPI_NONE, nullptr))) return rc;
    float* cur = b0;
    float* o1 = b1;
    float* o2 = b2;
    for (const sopro_mimi::Stage& S : m->stages) {
      if (Tn * S.ratio > 0x7fffffffLL) return mfail(SOPRO_ERR_INVALID, "sequence too long for one launch");
      // ELU -> ConvTranspose(stride r, kernel 2r) as a 2-tap implicit GEMM with r*Cout columns
      if ((rc = conv(cur, Tn, S.cin, 2, 1, Wd + S.tw, Wd + S.tb, S.ratio * S.cout, S.cout, o1, 1, EPI_NONE, nullptr))) return rc;
      Tn *= S.ratio;
      // ResnetBlock: o1 + conv1(ELU(conv3(ELU(o1))))
      const int hid = S.cout / c.compress;
      if ((rc = conv(o1, Tn, S.cout, c.res_kernel, c.res_kernel - 1, Wd + S.r1w, Wd + S.r1b, hid, hid, o2, 1, EPI_NONE, nullptr))) return rc;
      if ((rc = conv(o2, Tn, hid, 1, 0, Wd + S.r2w, Wd + S.r2b, S.cout, S.cout, cur, 1, EPI_RES, o1))) return rc;
      ch = S.cout;
    }
    final_conv_kernel<<<dim3((unsigned)((Tn + 255) / 256), B), 256, 0, st>>>(cur, Wd + m->lw, Wd + m->lb, wav, Tn, ch, c.last_kernel, 0);
    MCK(cudaGetLastError());
    return SOPRO_OK;
  }
  // tensor-core mode.  Activations that feed a contraction travel as bf16 with the consumer's ELU already
  // applied; only the ConvTranspose output (the ResnetBlock skip) and whatever a fp32 kernel reads are fp32.
  //   curh: bf16 operand of the next ConvTranspose;  z (b1): fp32 skip;  b0: fp32 block output when needed
  __nv_bfloat16* curh = h0;
  __nv_bfloat16* ha = h1;
  __nv_bfloat16* hb = h2;
  float* cur32 = nullptr;  // set when the running activation lives in fp32 (b0) instead of curh
  if (tc::supported(ch, c.kernel * C, C)) {
    const long long n4 = (long long)B * T2 * C / 4;
    cast_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x, hb, n4);
    MCK(cudaGetLastError());
    if ((rc = tcg(hb, Tn, C, c.kernel, c.kernel - 1, Wh + m->c0w_h, Wd + m->c0b, ch, ch, tc::EPI_NONE, nullptr, nullptr, nullptr, curh, 1)))
      return rc;
  } else {
    if ((rc = conv(x, Tn, C, c.kernel, c.kernel - 1, Wd + m->c0w, Wd + m->c0b, ch, ch, b0, 0, EPI_NONE, nullptr))) return rc;
    cur32 = b0;
  }
  const bool final_h = c.num_filters % 8 == 0 && c.last_kernel <= 8;  // final conv reads the bf16 ELU'd activation
  for (size_t si = 0; si < m->stages.size(); ++si) {
    const sopro_mimi::Stage& S = m->stages[si];
    if (Tn * S.ratio > 0x7fffffffLL) return mfail(SOPRO_ERR_INVALID, "sequence too long for one launch");
    const int hid = S.cout / c.compress, NT = S.ratio * S.cout;
    const bool last = si + 1 == m->stages.size();
    const bool t_ok = !cur32 && tc::supported(NT, 2 * S.cin, S.cin);
    const bool r1_ok = tc::supported(hid, c.res_kernel * S.cout, S.cout);
    const bool r2_ok = tc::supported(S.cout, hid, hid);
    // the consumer of this stage's output: the next ConvTranspose on tensor cores wants bf16 ELU(x); the
    // final conv and the fp32 kernels read fp32
    bool next_tc = final_h;  // after the last stage: the bf16 final conv
    if (!last) {
      const sopro_mimi::Stage& Nx = m->stages[si + 1];
      next_tc = tc::supported(Nx.ratio * Nx.cout, 2 * Nx.cin, Nx.cin);
    }
    // ConvTranspose -> z fp32 (b1) [+ bf16 ELU(z) in ha when res1 runs on tensor cores]
    if (t_ok) {
      if ((rc = tcg(curh, Tn, S.cin, 2, 1, Wh + S.tw_h, Wd + S.tb, NT, S.cout, tc::EPI_NONE, nullptr, nullptr, b1, r1_ok ? ha : nullptr, 1)))
        return rc;
    } else {
      if (!cur32) return mfail(SOPRO_ERR_INVALID, "internal: stage %zu has no fp32 input", si);
      if ((rc = conv(cur32, Tn, S.cin, 2, 1, Wd + S.tw, Wd + S.tb, NT, S.cout, b1, 1, EPI_NONE, nullptr))) return rc;
    }
    Tn *= S.ratio;
    const bool ha_valid = t_ok && r1_ok;
    // ResnetBlock in one launch when the geometry allows (hidden activation stays on chip), else conv by conv
    static const bool fuse_res = !(getenv("SOPRO_MIMI_FUSE_RES") && atoi(getenv("SOPRO_MIMI_FUSE_RES")) == 0);
    if (fuse_res && ha_valid && r2_ok && tc::resblock_supported(hid, S.cout) && (S.cout * c.res_kernel) % 64 == 0) {
      tc::ResOp ro{};
      ro.bias1 = Wd + S.r1b;
      ro.bias2 = Wd + S.r2b;
      ro.Z = b1;
      ro.out_f32 = next_tc ? nullptr : b0;
      ro.out_bf16 = next_tc ? curh : nullptr;
      ro.M = (int)Tn;
      ro.taps = c.res_kernel;
      ro.pad = c.res_kernel - 1;
      ro.out_elu = 1;
      cudaError_t fe = tc::launch_resblock(ha, Wh + S.r1w_h, Wh + S.r2w_h, hid, ro, B, st);
      if (fe != cudaSuccess) return mfail(SOPRO_ERR_CUDA, "fused ResnetBlock (stage %zu): %s", si, cudaGetErrorString(fe));
      cur32 = next_tc ? nullptr : b0;
      ch = S.cout;
      continue;
    }
    // res conv k=3 -> ELU(h) bf16 (hb) for a tensor-core res2, else raw h fp32 (b2)
    if (ha_valid) {
      if ((rc = tcg(ha, Tn, S.cout, c.res_kernel, c.res_kernel - 1, Wh + S.r1w_h, Wd + S.r1b, hid, hid, tc::EPI_NONE, nullptr, nullptr,
                    r2_ok ? nullptr : b2, r2_ok ? hb : nullptr, 1)))
        return rc;
    } else {
      if ((rc = conv(b1, Tn, S.cout, c.res_kernel, c.res_kernel - 1, Wd + S.r1w, Wd + S.r1b, hid, hid, b2, 1, EPI_NONE, nullptr))) return rc;
    }
    // res conv k=1 + skip
    if (ha_valid && r2_ok) {
      if ((rc = tcg(hb, Tn, hid, 1, 0, Wh + S.r2w_h, Wd + S.r2b, S.cout, S.cout, tc::EPI_RES, b1, nullptr, next_tc ? nullptr : b0,
                    next_tc ? curh : nullptr, 1)))
        return rc;
      cur32 = next_tc ? nullptr : b0;
    } else {
      if ((rc = conv(b2, Tn, hid, 1, 0, Wd + S.r2w, Wd + S.r2b, S.cout, S.cout, b0, 1, EPI_RES, b1))) return rc;
      cur32 = b0;
      if (next_tc && !last) {  // fp32 block output feeding a tensor-core ConvTranspose: not reachable with Mimi's geometry
        return mfail(SOPRO_ERR_UNSUPPORTED, "unsupported channel geometry for tensor-core mode (stage %zu)", si);
      }
    }
    ch = S.cout;
  }
  if (cur32) {
    final_conv_kernel<<<dim3((unsigned)((Tn + 255) / 256), B), 256, 0, st>>>(cur32, Wd + m->lw, Wd + m->lb, wav, Tn, ch, c.last_kernel, 0);
  } else {
    const int per = 256 - (c.last_kernel - 1);
    final_conv_h_kernel<<<dim3((unsigned)((Tn + per - 1) / per), B), 256, (size_t)(c.last_kernel * 256 + c.last_kernel * ch) * 4, st>>>(
        curh, Wd + m->lw, Wd + m->lb, wav, Tn, ch, c.last_kernel, 0);
  }
  MCK(cudaGetLastError());
  return SOPRO_OK;
}
}  // namespace


// ---------------------------------------------------------------------------------------------
// Streaming decode with persistent state (reference codec/mimi.py:83-181 MimiStreamDecoder; transformers
// modeling_mimi.py:77-170 MimiConv1dPaddingCache is the per-conv left context this replaces).
//
// A stream owns (a) one K/V ring per transformer layer holding the rotated keys and the values of the last
// `window` positions, (b) the previous RVQ frame (the depthwise ConvTranspose upsampler reads x[t-1]) and (c) for every
// causal conv of the SEANet decoder the last (taps-1) input rows.  (c) is stored IN PLACE: each conv input buffer is
// laid out [context rows | rows of this chunk]; the conv runs as a "valid" convolution over it (pad = 0,
// Min = M + taps - 1) and afterwards the buffer's last context-many rows are moved to its front.  At the start of a
// stream the context rows are zero, which is exactly the causal zero padding of the full decode, and every kernel
// computes each output element in the same order as the full decode: in fp32 mode the chunks are bit-identical to the
// full decode's prefix.  Work per chunk is O(chunk), not O(prefix).
// ---------------------------------------------------------------------------------------------
struct TailShift {
  void* base[16];
  int row_bytes[16], ctx[16], rows[16];  // rows = new rows written behind the ctx rows this step
  int n;
};

// one block per buffer: rows [rows, rows + ctx) -> [0, ctx) (through shared memory: the ranges may overlap)
__global__ void __launch_bounds__(256) tail_shift_kernel(const TailShift ts) {
  extern __shared__ uint4 tsm[];
  const int i = blockIdx.x;
  const int n16 = ts.ctx[i] * ts.row_bytes[i] / 16;
  const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(ts.base[i]) + (size_t)ts.rows[i] * ts.row_bytes[i]);
  uint4* dst = reinterpret_cast<uint4*>(ts.base[i]);
  for (int e = threadIdx.x; e < n16; e += blockDim.x) tsm[e] = src[e];
  __syncthreads();
  for (int e = threadIdx.x; e < n16; e += blockDim.x) dst[e] = tsm[e];
}

struct sopro_mimi_stream {
  sopro_mimi* m = nullptr;
  int max_n = 0, precision = 0, R = 0;
  long long frames = 0;
  unsigned char* slab = nullptr;
  size_t slab_bytes = 0, state_bytes = 0;
  // ---- state (zeroed by reset): [up_prev | K rings | V rings | conv-context rows at the front of the buffers below]
  float* up_prev = nullptr;
  float *kring = nullptr, *vring = nullptr;  // [n_layers][R][C]
  // ---- chunk buffers
  float *S = nullptr, *E = nullptr, *XC = nullptr, *LN = nullptr, *QKV = nullptr, *ATT = nullptr, *HID = nullptr;
  __nv_bfloat16 *LNh = nullptr, *ATTh = nullptr, *HIDh = nullptr, *XCh = nullptr;
  void* A0 = nullptr;               // conv0 output  [1 + T2][16F]    (fp32 raw | bf16 ELU'd)
  void* Z[SOPRO_MIMI_MAX_RATIOS]{};    // ConvTranspose output [2 + Tn][cout] (fp32 raw | bf16 ELU'd)
  float* Zf[SOPRO_MIMI_MAX_RATIOS]{};  // tensor-core mode: fp32 skip [Tn][cout]
  float* Hs[SOPRO_MIMI_MAX_RATIOS]{};  // fp32 mode: res hidden [Tn][cout/2]
  void* O[SOPRO_MIMI_MAX_RATIOS]{};    // block output [ctx + Tn][cout], ctx = 1 (next ConvTranspose) or taps-1 (final conv)
  int* codes_dev = nullptr;
  float* wav_dev = nullptr;
};

namespace {
struct SlabPlan {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off = (off + bytes + 255) / 256 * 256;
    return o;
  }
};

// lays the stream's slab out; with base == nullptr only sizes are computed
void stream_layout(sopro_mimi_stream* s, unsigned char* base) {
  const sopro_mimi_config_t& c = s->m->cfg;
  const bool tcm = s->precision == SOPRO_MIMI_BF16_TC;
  const size_t C = c.hidden, FF = c.ffn, n = s->max_n, T2 = 2 * n, NL = c.n_layers;
  const size_t es = tcm ? 2 : 4;  // element size of the conv operands
  SlabPlan P;
  auto at = [&](size_t o) { return base ? base + o : nullptr; };
  // state first
  s->up_prev = (float*)at(P.take(C * 4));
  s->kring = (float*)at(P.take(NL * s->R * C * 4));
  s->vring = (float*)at(P.take(NL * s->R * C * 4));
  // conv operand buffers: the context rows at their fronts are state too, so they come next
  const int k0 = c.kernel - 1;
  s->XC = (float*)at(P.take((k0 + T2) * C * 4));
  s->XCh = tcm ? (__nv_bfloat16*)at(P.take((k0 + T2) * C * 2)) : nullptr;
  size_t ch = (size_t)c.num_filters << c.n_ratios, Tn = T2;
  s->A0 = at(P.take((1 + Tn) * ch * es));
  for (int i = 0; i < c.n_ratios; ++i) {
    const size_t cout = ch / 2;
    Tn *= c.ratios[i];
    const size_t ctx_o = i + 1 == c.n_ratios ? (size_t)(c.last_kernel - 1) : 1;
    s->Z[i] = at(P.take((c.res_kernel - 1 + Tn) * cout * es));
    s->O[i] = at(P.take((ctx_o + Tn) * cout * es));
    ch = cout;
  }
  s->state_bytes = P.off;  // everything up to here is zeroed by reset (a superset of the state proper)
  ch = (size_t)c.num_filters << c.n_ratios;
  Tn = T2;
  for (int i = 0; i < c.n_ratios; ++i) {
    const size_t cout = ch / 2;
    Tn *= c.ratios[i];
    s->Zf[i] = tcm ? (float*)at(P.take(Tn * cout * 4)) : nullptr;
    s->Hs[i] = (float*)at(P.take(Tn * (cout / c.compress) * 4));  // fp32 mode: fp32; tensor-core mode: bf16 view (unfused blocks)
    ch = cout;
  }
  s->S = (float*)at(P.take(n * C * 4));
  s->E = (float*)at(P.take(n * C * 4));
  s->LN = (float*)at(P.take(T2 * C * 4));
  s->QKV = (float*)at(P.take(T2 * 3 * C * 4));
  s->ATT = (float*)at(P.take(T2 * C * 4));
  s->HID = (float*)at(P.take(T2 * FF * 4));
  s->LNh = (__nv_bfloat16*)s->LN;
  s->ATTh = (__nv_bfloat16*)s->ATT;
  s->HIDh = (__nv_bfloat16*)s->HID;
  s->slab_bytes = P.off;
}

int stream_step(sopro_mimi_stream* s, const int32_t* codes, int n, int code_stride, float* wav, cudaStream_t st) {
  sopro_mimi* m = s->m;
  const sopro_mimi_config_t& c = m->cfg;
  const int C = c.hidden, T2 = 2 * n, H = c.n_heads, Dh = C / H, FF = c.ffn;
  const float* Wd = m->dev;
  const __nv_bfloat16* Wh = m->dev_h;
  const bool tcm = s->precision == SOPRO_MIMI_BF16_TC;
  const int pos0 = (int)(2 * s->frames);
  int rc = ensure_rope(m, std::max(4096, 2 * (pos0 + T2)), st);
  if (rc) return rc;
  GemmOp g{};
  auto lin = [&](const float* A, int M, int K, const float* W, int N, float* Cc, int epi, const float* R, const float* scale) {
    g = GemmOp{};
    g.A = A; g.W = W; g.C = Cc; g.R = R; g.scale = scale; g.bias = nullptr;
    g.M = M; g.N = N; g.K = K; g.Min = M; g.Cin = K; g.taps = 1; g.dil = 1; g.pad = 0; g.ldc = N; g.bias_mod = N; g.epi = epi;
    return launch_gemm(g, 1, st);
  };
  // "valid" conv over [ctx rows | M rows]: A points at the first context row, Min = M + taps - 1, pad = 0
  auto conv = [&](const float* A, long long M, int cin, int taps, const float* W, const float* bias, int N, int bias_mod, float* Cc,
                  int elu, int epi, const float* R) {
    g = GemmOp{};
    g.A = A; g.W = W; g.C = Cc; g.R = R; g.bias = bias; g.scale = nullptr;
    g.M = (int)M; g.N = N; g.K = taps * cin; g.Min = (int)M + taps - 1; g.Cin = cin; g.taps = taps; g.dil = 1; g.pad = 0; g.ldc = N;
    g.bias_mod = bias_mod; g.epi = epi; g.a_elu = elu;
    return launch_gemm(g, 1, st);
  };
  auto tcg = [&](const __nv_bfloat16* X, long long M, int cin, int taps, const __nv_bfloat16* W, const float* bias, int N, int bias_mod,
                 int epi, const float* R, const float* scale, float* of, __nv_bfloat16* oh, int out_elu) {
    tc::TcOp o{};
    o.bias = bias; o.R = R; o.scale = scale; o.out_f32 = of; o.out_bf16 = oh;
    o.c_bs = M * N; o.M = (int)M; o.N = N; o.K = taps * cin; o.Cin = cin; o.dil = 1; o.pad = 0;
    o.bias_mod = bias_mod; o.epi = epi; o.out_elu = out_elu;
    cudaError_t e = tc::launch(X, M + taps - 1, W, o, 1, st);
    if (e != cudaSuccess) return mfail(SOPRO_ERR_CUDA, "tensor-core GEMM (stream, N=%d K=%d): %s", N, o.K, cudaGetErrorString(e));
    return (int)SOPRO_OK;
  };
  // ---- RVQ + projection + upsample
  rvq_gather_kernel<<<dim3(n, 1), 256, 0, st>>>(codes, Wd + m->embed, s->S, c.n_q, code_stride, c.codebook_dim, c.vocab, c.n_sem, m->bad_code);
  MCK(cudaGetLastError());
  if ((rc = lin(s->S, n, C, Wd + m->rvq_w, C, s->E, EPI_NONE, nullptr, nullptr))) return rc;
  const int k0 = c.kernel - 1;
  float* x = s->XC + (size_t)k0 * C;  // residual stream: the rows behind conv0's context rows
  upsample_kernel<<<dim3(T2, 1), 256, 0, st>>>(s->E, Wd + m->up_w, x, n, C, s->up_prev);
  MCK(cudaGetLastError());
  MCK(cudaMemcpyAsync(s->up_prev, s->E + (size_t)(n - 1) * C, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
  // ---- transformer: K/V of the new positions go to the rings, queries attend over the ring
  const unsigned ln_grid = (unsigned)((T2 + 7) / 8);
  const size_t asm_bytes = (size_t)8 * (Dh + c.window) * 4;
  for (size_t li = 0; li < m->layers.size(); ++li) {
    const sopro_mimi::Layer& L = m->layers[li];
    float* kr = s->kring + li * (size_t)s->R * C;
    float* vr = s->vring + li * (size_t)s->R * C;
    if (tcm) {
      layernorm_kernel<<<ln_grid, 256, 0, st>>>(x, Wd + L.ln1w, Wd + L.ln1b, s->LNh, (long long)T2, C, c.norm_eps);
      if ((rc = tcg(s->LNh, T2, C, 1, Wh + L.qkv_h, nullptr, 3 * C, 3 * C, tc::EPI_NONE, nullptr, nullptr, s->QKV, nullptr, 0))) return rc;
      rope_kernel<<<dim3(T2, 1), 256, 0, st>>>(s->QKV, m->rope, T2, m->rope_T2, C, H, pos0, kr, vr, s->R);
      attn_kernel<<<dim3((T2 + 7) / 8, H, 1), 256, asm_bytes, st>>>(s->QKV, s->ATTh, T2, C, H, c.window, pos0, kr, vr, s->R);
      MCK(cudaGetLastError());
      if ((rc = tcg(s->ATTh, T2, C, 1, Wh + L.wo_h, nullptr, C, C, tc::EPI_RES_SCALE, x, Wd + L.ls1, x, nullptr, 0))) return rc;
      layernorm_kernel<<<ln_grid, 256, 0, st>>>(x, Wd + L.ln2w, Wd + L.ln2b, s->LNh, (long long)T2, C, c.norm_eps);
      if ((rc = tcg(s->LNh, T2, C, 1, Wh + L.fc1_h, nullptr, FF, FF, tc::EPI_GELU, nullptr, nullptr, nullptr, s->HIDh, 0))) return rc;
      if ((rc = tcg(s->HIDh, T2, FF, 1, Wh + L.fc2_h, nullptr, C, C, tc::EPI_RES_SCALE, x, Wd + L.ls2, x, nullptr, 0))) return rc;
    } else {
      layernorm_kernel<<<ln_grid, 256, 0, st>>>(x, Wd + L.ln1w, Wd + L.ln1b, s->LN, (long long)T2, C, c.norm_eps);
      if ((rc = lin(s->LN, T2, C, Wd + L.qkv, 3 * C, s->QKV, EPI_NONE, nullptr, nullptr))) return rc;
      rope_kernel<<<dim3(T2, 1), 256, 0, st>>>(s->QKV, m->rope, T2, m->rope_T2, C, H, pos0, kr, vr, s->R);
      attn_kernel<<<dim3((T2 + 7) / 8, H, 1), 256, asm_bytes, st>>>(s->QKV, s->ATT, T2, C, H, c.window, pos0, kr, vr, s->R);
      MCK(cudaGetLastError());
      if ((rc = lin(s->ATT, T2, C, Wd + L.wo, C, x, EPI_RES_SCALE, x, Wd + L.ls1))) return rc;
      layernorm_kernel<<<ln_grid, 256, 0, st>>>(x, Wd + L.ln2w, Wd + L.ln2b, s->LN, (long long)T2, C, c.norm_eps);
      if ((rc = lin(s->LN, T2, C, Wd + L.fc1, FF, s->HID, EPI_GELU, nullptr, nullptr))) return rc;
      if ((rc = lin(s->HID, T2, FF, Wd + L.fc2, C, x, EPI_RES_SCALE, x, Wd + L.ls2))) return rc;
    }
  }
  // ---- SEANet decoder over [context | chunk] buffers
  TailShift ts{};
  auto carry = [&](void* base, int row_bytes, int ctx, long long rows) {
    ts.base[ts.n] = base;
    ts.row_bytes[ts.n] = row_bytes;
    ts.ctx[ts.n] = ctx;
    ts.rows[ts.n] = (int)rows;
    ++ts.n;
  };
  long long Tn = T2;
  int ch = c.num_filters << c.n_ratios;
  const int kr3 = c.res_kernel - 1;
  if (!tcm) {
    float* a0 = reinterpret_cast<float*>(s->A0);
    if ((rc = conv(s->XC, Tn, C, c.kernel, Wd + m->c0w, Wd + m->c0b, ch, ch, a0 + (size_t)ch, 0, EPI_NONE, nullptr))) return rc;
    carry(s->XC, C * 4, k0, Tn);
    carry(a0, ch * 4, 1, Tn);
    const float* cur = a0;  // [1 ctx row | Tn rows]
    for (size_t si = 0; si < m->stages.size(); ++si) {
      const sopro_mimi::Stage& S = m->stages[si];
      const bool last = si + 1 == m->stages.size();
      const int hid = S.cout / c.compress, ctx_o = last ? c.last_kernel - 1 : 1;
      float* z = reinterpret_cast<float*>(s->Z[si]);
      float* o = reinterpret_cast<float*>(s->O[si]);
      if ((rc = conv(cur, Tn, S.cin, 2, Wd + S.tw, Wd + S.tb, S.ratio * S.cout, S.cout, z + (size_t)kr3 * S.cout, 1, EPI_NONE, nullptr))) return rc;
      Tn *= S.ratio;
      if ((rc = conv(z, Tn, S.cout, c.res_kernel, Wd + S.r1w, Wd + S.r1b, hid, hid, s->Hs[si], 1, EPI_NONE, nullptr))) return rc;
      if ((rc = conv(s->Hs[si], Tn, hid, 1, Wd + S.r2w, Wd + S.r2b, S.cout, S.cout, o + (size_t)ctx_o * S.cout, 1, EPI_RES,
                     z + (size_t)kr3 * S.cout)))
        return rc;
      carry(z, S.cout * 4, kr3, Tn);
      carry(o, S.cout * 4, ctx_o, Tn);
      cur = o;
      ch = S.cout;
    }
    const int lk = c.last_kernel - 1;
    final_conv_kernel<<<dim3((unsigned)((Tn + 255) / 256), 1), 256, 0, st>>>(cur + (size_t)lk * ch, Wd + m->lw, Wd + m->lb, wav, Tn, ch,
                                                                              c.last_kernel, -lk);
    MCK(cudaGetLastError());
  } else {
    if (!tc::supported(ch, c.kernel * C, C) || c.num_filters % 8 || c.last_kernel > 8)
      return mfail(SOPRO_ERR_UNSUPPORTED, "streaming tensor-core mode: unsupported conv0 / final conv geometry (use SOPRO_MIMI_FP32)");
    const long long n4 = (long long)T2 * C / 4;
    cast_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x, s->XCh + (size_t)k0 * C, n4);
    MCK(cudaGetLastError());
    __nv_bfloat16* a0 = reinterpret_cast<__nv_bfloat16*>(s->A0);
    if ((rc = tcg(s->XCh, Tn, C, c.kernel, Wh + m->c0w_h, Wd + m->c0b, ch, ch, tc::EPI_NONE, nullptr, nullptr, nullptr, a0 + (size_t)ch, 1)))
      return rc;
    carry(s->XCh, C * 2, k0, Tn);
    carry(a0, ch * 2, 1, Tn);
    const __nv_bfloat16* cur = a0;
    for (size_t si = 0; si < m->stages.size(); ++si) {
      const sopro_mimi::Stage& S = m->stages[si];
      const bool last = si + 1 == m->stages.size();
      const int hid = S.cout / c.compress, NT = S.ratio * S.cout, ctx_o = last ? c.last_kernel - 1 : 1;
      if (!tc::supported(NT, 2 * S.cin, S.cin) || !tc::supported(hid, c.res_kernel * S.cout, S.cout) || !tc::supported(S.cout, hid, hid))
        return mfail(SOPRO_ERR_UNSUPPORTED, "streaming tensor-core mode: unsupported geometry at stage %zu (use SOPRO_MIMI_FP32)", si);
      const bool fused = tc::resblock_supported(hid, S.cout) && (S.cout * c.res_kernel) % 64 == 0;
      __nv_bfloat16* zh = reinterpret_cast<__nv_bfloat16*>(s->Z[si]);
      __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(s->O[si]);
      if ((rc = tcg(cur, Tn, S.cin, 2, Wh + S.tw_h, Wd + S.tb, NT, S.cout, tc::EPI_NONE, nullptr, nullptr, s->Zf[si], zh + (size_t)kr3 * S.cout, 1)))
        return rc;
      Tn *= S.ratio;
      if (fused) {
        tc::ResOp ro{};
        ro.bias1 = Wd + S.r1b;
        ro.bias2 = Wd + S.r2b;
        ro.Z = s->Zf[si];
        ro.out_f32 = nullptr;
        ro.out_bf16 = oh + (size_t)ctx_o * S.cout;
        ro.M = (int)Tn;
        ro.Min = (int)Tn + kr3;
        ro.taps = c.res_kernel;
        ro.pad = 0;
        ro.out_elu = 1;
        cudaError_t fe = tc::launch_resblock(zh, Wh + S.r1w_h, Wh + S.r2w_h, hid, ro, 1, st);
        if (fe != cudaSuccess) return mfail(SOPRO_ERR_CUDA, "fused ResnetBlock (stream, stage %zu): %s", si, cudaGetErrorString(fe));
      } else {  // conv k=3 -> ELU(h) bf16, then the 1x1 conv + fp32 skip (same two launches as the one-shot decode)
        __nv_bfloat16* hh = reinterpret_cast<__nv_bfloat16*>(s->Hs[si]);
        if ((rc = tcg(zh, Tn, S.cout, c.res_kernel, Wh + S.r1w_h, Wd + S.r1b, hid, hid, tc::EPI_NONE, nullptr, nullptr, nullptr, hh, 1))) return rc;
        if ((rc = tcg(hh, Tn, hid, 1, Wh + S.r2w_h, Wd + S.r2b, S.cout, S.cout, tc::EPI_RES, s->Zf[si], nullptr, nullptr,
                      oh + (size_t)ctx_o * S.cout, 1)))
          return rc;
      }
      carry(zh, S.cout * 2, kr3, Tn);
      carry(oh, S.cout * 2, ctx_o, Tn);
      cur = oh;
      ch = S.cout;
    }
    const int lk = c.last_kernel - 1, per = 256 - lk;
    final_conv_h_kernel<<<dim3((unsigned)((Tn + per - 1) / per), 1), 256, (size_t)(c.last_kernel * 256 + c.last_kernel * ch) * 4, st>>>(
        cur + (size_t)lk * ch, Wd + m->lw, Wd + m->lb, wav, Tn, ch, c.last_kernel, -lk);
    MCK(cudaGetLastError());
  }
  tail_shift_kernel<<<ts.n, 256, 16384, st>>>(ts);
  MCK(cudaGetLastError());
  s->frames += n;
  return SOPRO_OK;
}
}  // namespace

extern "C" {

int sopro_mimi_stream_create(sopro_mimi_t* m, int max_chunk_frames, sopro_mimi_stream_t** out) {
  if (!m || !out) return mfail(SOPRO_ERR_INVALID, "null argument");
  *out = nullptr;
  if (max_chunk_frames < 1 || max_chunk_frames > 256) return mfail(SOPRO_ERR_INVALID, "max_chunk_frames must be in [1, 256]");
  MCK(cudaSetDevice(m->device));
  sopro_mimi_stream* s = new sopro_mimi_stream();
  s->m = m;
  s->max_n = max_chunk_frames;
  s->precision = m->precision;
  s->R = (m->cfg.window + 2 * max_chunk_frames + 7) / 8 * 8;
  stream_layout(s, nullptr);
  if ((size_t)(m->cfg.kernel - 1) * m->cfg.hidden * 4 > 16384) {
    delete s;
    return mfail(SOPRO_ERR_UNSUPPORTED, "conv context rows exceed the tail-shift staging buffer");
  }
  cudaError_t e = cudaMalloc(&s->slab, s->slab_bytes);
  if (e != cudaSuccess) {
    delete s;
    return mfail(SOPRO_ERR_CUDA, "stream state %zu MB: %s", s->slab_bytes >> 20, cudaGetErrorString(e));
  }
  stream_layout(s, s->slab);
  e = cudaMemset(s->slab, 0, s->state_bytes);
  if (e != cudaSuccess) {
    cudaFree(s->slab);
    delete s;
    return mfail(SOPRO_ERR_CUDA, "stream state init: %s", cudaGetErrorString(e));
  }
  *out = s;
  return SOPRO_OK;
}

int sopro_mimi_stream_destroy(sopro_mimi_stream_t* s) {
  if (!s) return SOPRO_OK;
  cudaSetDevice(s->m->device);
  cudaFree(s->slab);
  cudaFree(s->codes_dev);
  delete s;
  return SOPRO_OK;
}

int sopro_mimi_stream_reset(sopro_mimi_stream_t* s, void* stream) {
  if (!s) return mfail(SOPRO_ERR_INVALID, "null argument");
  MCK(cudaSetDevice(s->m->device));
  if (s->precision != s->m->precision) {  // the buffers are laid out per arithmetic mode
    s->precision = s->m->precision;
    size_t old = s->slab_bytes;
    stream_layout(s, nullptr);
    if (s->slab_bytes > old) {
      MCK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
      cudaFree(s->slab);
      s->slab = nullptr;
      MCK(cudaMalloc(&s->slab, s->slab_bytes));
    }
    stream_layout(s, s->slab);
  }
  MCK(cudaMemsetAsync(s->slab, 0, s->state_bytes, reinterpret_cast<cudaStream_t>(stream)));
  s->frames = 0;
  return SOPRO_OK;
}

int64_t sopro_mimi_stream_frames(const sopro_mimi_stream_t* s) { return s ? s->frames : -1; }

int sopro_mimi_decode_step(sopro_mimi_stream_t* s, const int32_t* codes, int n, float* wav, void* stream) {
  if (!s || !codes || !wav) return mfail(SOPRO_ERR_INVALID, "null argument");
  if (n < 1) return mfail(SOPRO_ERR_INVALID, "n must be >= 1");
  if (s->precision != s->m->precision)
    return mfail(SOPRO_ERR_STATE, "the decoder's precision changed since this stream started: call sopro_mimi_stream_reset");
  if (2 * (s->frames + n) > 0x3fffffffLL) return mfail(SOPRO_ERR_INVALID, "stream too long");
  MCK(cudaSetDevice(s->m->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t hop = sopro_mimi_samples_per_frame(s->m);
  for (int done = 0; done < n; done += s->max_n) {  // codes are [n_q][n]: a sub-chunk starts at column `done`
    const int k = std::min(s->max_n, n - done);
    const int rc = stream_step(s, codes + done, k, n, wav + (size_t)done * hop, st);
    if (rc) return rc;
  }
  return SOPRO_OK;
}

int sopro_mimi_decode_step_host(sopro_mimi_stream_t* s, const int32_t* codes_host, int n, float* wav_host, void* stream) {
  if (!s || !codes_host || !wav_host) return mfail(SOPRO_ERR_INVALID, "null argument");
  if (n < 1 || n > 65536) return mfail(SOPRO_ERR_INVALID, "n must be in [1, 65536]");
  const sopro_mimi_config_t& c = s->m->cfg;
  for (size_t i = 0; i < (size_t)n * c.n_q; ++i)
    if (codes_host[i] < 0 || codes_host[i] >= c.vocab)
      return mfail(SOPRO_ERR_INVALID, "code %d at flat index %zu is outside [0, %d)", codes_host[i], i, c.vocab);
  MCK(cudaSetDevice(s->m->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t nc = (size_t)n * c.n_q, nw = (size_t)n * (size_t)sopro_mimi_samples_per_frame(s->m);
  cudaFree(s->codes_dev);
  s->codes_dev = nullptr;
  MCK(cudaMalloc(&s->codes_dev, nc * 4 + nw * 4));
  s->wav_dev = reinterpret_cast<float*>(s->codes_dev + nc);
  MCK(cudaMemcpyAsync(s->codes_dev, codes_host, nc * 4, cudaMemcpyHostToDevice, st));
  const int rc = sopro_mimi_decode_step(s, s->codes_dev, n, s->wav_dev, stream);
  if (rc) return rc;
  MCK(cudaMemcpyAsync(wav_host, s->wav_dev, nw * 4, cudaMemcpyDeviceToHost, st));
  MCK(cudaStreamSynchronize(st));
  return SOPRO_OK;
}

}  // extern "C"

extern "C" {

int sopro_mimi_check(sopro_mimi_t* m, void* stream) {
  if (!m) return mfail(SOPRO_ERR_INVALID, "null argument");
  MCK(cudaSetDevice(m->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int flag = 0;
  MCK(cudaMemcpyAsync(&flag, m->bad_code, 4, cudaMemcpyDeviceToHost, st));
  MCK(cudaMemsetAsync(m->bad_code, 0, 4, st));
  MCK(cudaStreamSynchronize(st));
  if (flag) return mfail(SOPRO_ERR_INVALID, "a decode since the last check read a code outside [0, %d) (clamped)", m->cfg.vocab);
  return SOPRO_OK;
}

int sopro_mimi_set_graphs(sopro_mimi_t* m, int enabled) {
  if (!m) return mfail(SOPRO_ERR_INVALID, "null argument");
  m->graphs = enabled != 0;
  return SOPRO_OK;
}

int sopro_debug_tc_gemm(const void* X, int B, int64_t rows, int cin, int taps, int dil, int pad, const void* W, int N,
                        const float* bias, int bias_mod, int epi, const float* R, const float* scale, float* out_f32,
                        void* out_bf16, int out_elu, void* stream) {
  if (!X || !W || (!out_f32 && !out_bf16)) return mfail(SOPRO_ERR_INVALID, "null argument");
  if (B < 1 || B > 65535 || rows < 1 || rows > 0x7fffffffLL || !tc::supported(N, taps * cin, cin))
    return mfail(SOPRO_ERR_INVALID, "tc gemm: unsupported shape (rows=%lld cin=%d taps=%d N=%d)", (long long)rows, cin, taps, N);
  tc::TcOp o{};
  o.bias = bias; o.R = R; o.scale = scale; o.out_f32 = out_f32; o.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  o.c_bs = rows * N; o.M = (int)rows; o.N = N; o.K = taps * cin; o.Cin = cin; o.dil = dil; o.pad = pad;
  o.bias_mod = bias_mod > 0 ? bias_mod : N; o.epi = epi; o.out_elu = out_elu;
  cudaError_t e = tc::launch(X, rows, W, o, B, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return mfail(SOPRO_ERR_CUDA, "tensor-core GEMM launch: %s", cudaGetErrorString(e));
  return SOPRO_OK;
}

int sopro_mimi_decode_host(sopro_mimi_t* m, const int32_t* codes_host, int B, int T, float* wav_host, void* stream) {
  if (!m || !codes_host || !wav_host) return mfail(SOPRO_ERR_INVALID, "null argument");
  MCK(cudaSetDevice(m->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B < 1 || T < 1) return mfail(SOPRO_ERR_INVALID, "B and T must be >= 1");
  const size_t nc = (size_t)B * m->cfg.n_q * T;
  const size_t nw = (size_t)B * T * sopro_mimi_samples_per_frame(m);
  for (size_t i = 0; i < nc; ++i)
    if (codes_host[i] < 0 || codes_host[i] >= m->cfg.vocab)
      return mfail(SOPRO_ERR_INVALID, "code %d at flat index %zu is outside [0, %d)", codes_host[i], i, m->cfg.vocab);
  if (m->codes_cap < nc * 4 + nw * 4) {
    cudaFree(m->codes_dev);
    m->codes_dev = nullptr;
    m->codes_cap = 0;
    MCK(cudaMalloc(&m->codes_dev, nc * 4 + nw * 4));
    m->codes_cap = nc * 4 + nw * 4;
  }
  float* wav_dev = reinterpret_cast<float*>(m->codes_dev + nc);
  MCK(cudaMemcpyAsync(m->codes_dev, codes_host, nc * 4, cudaMemcpyHostToDevice, st));
  int rc = sopro_mimi_decode(m, m->codes_dev, B, T, wav_dev, stream);
  if (rc) return rc;
  MCK(cudaMemcpyAsync(wav_host, wav_dev, nw * 4, cudaMemcpyDeviceToHost, st));
  MCK(cudaStreamSynchronize(st));
  return SOPRO_OK;
}

}  // extern "C"

// =====================================================================================================================
// ENCODE: waveform -> codes (MimiModel.encode, modeling_mimi.py:1455-1488, 1522-1611; the reference calls it once per
// reference voice, codec/mimi.py:41-63).  fp32 throughout on the kernels above: every conv is the implicit GEMM; a conv
// of kernel 2r and stride r over [L][C] is the 2-tap stride-1 conv over the same memory read as [L/r][r*C] (L padded
// with zero rows to a multiple of r: MimiConv1d's "extra padding", :273-285; its causal left padding of r rows is the
// tap at superrow -1).
// =====================================================================================================================
namespace mimi {
// first conv: wav [L] -> y [L][F], kernel k, causal (left zero pad k-1), weight [F][k]
__global__ void enc_conv0_kernel(const float* __restrict__ wav, const float* __restrict__ w, const float* __restrict__ bias,
                                 float* __restrict__ y, long long L, int F, int k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * F) return;
  const long long t = i / F;
  const int c = (int)(i - t * F);
  float acc = __ldg(bias + c);
  for (int j = 0; j < k; ++j) {
    const long long ti = t + j - (k - 1);
    if (ti >= 0) acc = fmaf(__ldg(w + c * k + j), __ldg(wav + ti), acc);
  }
  y[i] = acc;
}

// replicate padding for the 25 -> 12.5 Hz conv: y rows [0, left) = x[0], [left, left+T) = x, [left+T, rows) = x[T-1]
__global__ void replicate_pad_kernel(const float* __restrict__ x, float* __restrict__ y, int T, int C, int left, int rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * C) return;
  const int r = (int)(i / C), c = (int)(i - (long long)r * C);
  int src = r - left;
  src = src < 0 ? 0 : (src >= T ? T - 1 : src);
  y[i] = x[(size_t)src * C + c];
}

// Residual nearest-neighbour search (MimiResidualVectorQuantizer.encode :1262-1280 with MimiEuclideanCodebook.quantize
// :1197-1203): one CTA per frame, the residual (Dc = 32*DPL floats) in registers, lane-sliced; a warp scans every 8th
// code vector, squared distance summed directly (the reference's cdist goes through |x|^2+|e|^2-2xe, same minimiser),
// lowest index wins ties (torch.argmin).  proj [T][2*Dc] = [semantic input_proj | acoustic input_proj] of the latent.
template <int DPL>
__global__ void __launch_bounds__(256) rvq_encode_kernel(const float* __restrict__ proj, const float* __restrict__ embed,
                                                         int* __restrict__ codes, int T, int n_q, int n_sem, int V) {
  constexpr int Dc = 32 * DPL;
  __shared__ float best_d[8];
  __shared__ int best_i[8];
  __shared__ int winner;
  const int t = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float r[DPL];
  for (int q = 0; q < n_q; ++q) {
    if (q == 0 || q == n_sem) {
      const float* p = proj + (size_t)t * 2 * Dc + (q == 0 ? 0 : Dc) + lane * DPL;
#pragma unroll
      for (int i = 0; i < DPL; ++i) r[i] = p[i];
    }
    const float* E = embed + (size_t)q * V * Dc;
    float bd = INFINITY;
    int bi = 0x7fffffff;
    for (int k = warp; k < V; k += 8) {
      const float* e = E + (size_t)k * Dc + lane * DPL;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < DPL; i += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(e + i));
        float d = r[i] - v.x;
        acc = fmaf(d, d, acc);
        d = r[i + 1] - v.y;
        acc = fmaf(d, d, acc);
        d = r[i + 2] - v.z;
        acc = fmaf(d, d, acc);
        d = r[i + 3] - v.w;
        acc = fmaf(d, d, acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (acc < bd) {  // k ascends within a warp: strict < keeps the lowest index
        bd = acc;
        bi = k;
      }
    }
    if (lane == 0) {
      best_d[warp] = bd;
      best_i[warp] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float d = best_d[0];
      int ix = best_i[0];
      for (int w = 1; w < 8; ++w)
        if (best_d[w] < d || (best_d[w] == d && best_i[w] < ix)) {
          d = best_d[w];
          ix = best_i[w];
        }
      winner = ix;
      codes[(size_t)q * T + t] = ix;
    }
    __syncthreads();
    const float* e = E + (size_t)winner * Dc + lane * DPL;
#pragma unroll
    for (int i = 0; i < DPL; ++i) r[i] -= __ldg(e + i);
    // (best_d / best_i / winner are rewritten only after the next __syncthreads pair)
  }
}
}  // namespace mimi

struct sopro_mimi_encoder {
  int device = 0;
  sopro_mimi_config_t cfg{};
  float* dev = nullptr;
  size_t c0w = 0, c0b = 0, lw = 0, lb = 0, down_w = 0, inproj = 0, embed = 0;
  struct Stage {
    size_t r1w, r1b, r2w, r2b, dw, db;
    int ratio, cin;
  };
  std::vector<Stage> stages;
  std::vector<sopro_mimi::Layer> layers;  // fp32 offsets only
  float* rope = nullptr;
  int rope_T2 = 0;
  float* ws = nullptr;
  size_t ws_bytes = 0;
  float* wav_dev = nullptr;   // staging of the *_host entry point
  int* codes_dev = nullptr;
  float* lat_dev = nullptr;
  size_t wav_cap = 0, codes_cap = 0, lat_cap = 0;
};

namespace {
constexpr long long kEncMaxSamples = 24000LL * 600;  // ten minutes of audio; a reference voice is seconds

struct EncPlan {
  long long len[SOPRO_MIMI_MAX_RATIOS + 1];   // rows entering stage s (len[0] = samples), len[n_ratios] = transformer positions
  long long padded[SOPRO_MIMI_MAX_RATIOS];    // len[s] rounded up to the stage's stride
  long long T;                                // frames
  size_t buf, xsz, need;
};

EncPlan enc_plan(const sopro_mimi_config_t& c, long long n) {
  EncPlan p{};
  p.len[0] = n;
  size_t widest = 0;
  int ch = c.num_filters;
  for (int s = 0; s < c.n_ratios; ++s) {
    const int r = c.ratios[c.n_ratios - 1 - s];
    p.padded[s] = (p.len[s] + r - 1) / r * r;
    p.len[s + 1] = p.padded[s] / r;
    widest = std::max(widest, (size_t)p.padded[s] * ch);
    ch *= 2;
  }
  const long long T2 = p.len[c.n_ratios];
  p.T = (T2 + 1) / 2;
  widest = std::max(widest, (size_t)T2 * ch);                                   // last conv input [T2][16F]
  widest = std::max(widest, (size_t)T2 * (size_t)std::max(3 * c.hidden, c.ffn));  // QKV / MLP hidden
  widest = std::max(widest, (size_t)(2 * p.T + 2) * c.hidden);                  // replicate-padded downsample input
  p.buf = (widest + 63) / 64 * 64;
  p.xsz = ((size_t)T2 * c.hidden + 63) / 64 * 64;
  p.need = (3 * p.buf + 2 * p.xsz) * 4;
  return p;
}
}  // namespace

extern "C" {

int sopro_mimi_encoder_create(const sopro_mimi_config_t* cfg, const sopro_mimi_encoder_weights_t* w, int device,
                              sopro_mimi_encoder_t** out) {
  if (!cfg || !w || !out) return mfail(SOPRO_ERR_INVALID, "null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0) return mfail(SOPRO_ERR_UNSUPPORTED, "no CUDA device; the Mimi encoder has no CPU fallback");
  if (device < 0 || device >= ndev) return mfail(SOPRO_ERR_INVALID, "device %d out of range", device);
  cudaDeviceProp prop;
  MCK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return mfail(SOPRO_ERR_UNSUPPORTED, "device is sm_%d%d; this build targets sm_100a only", prop.major, prop.minor);
  const int C = cfg->hidden, Dc = cfg->codebook_dim, Q = cfg->n_q, V = cfg->vocab, NL = cfg->n_layers, FF = cfg->ffn, F0 = cfg->num_filters;
  if (C % 64 || Dc != 256 || C != 2 * Dc || NL < 1 || NL > SOPRO_MIMI_MAX_LAYERS || cfg->n_ratios < 1 || cfg->n_ratios > SOPRO_MIMI_MAX_RATIOS ||
      cfg->n_heads < 1 || C % cfg->n_heads || (C / cfg->n_heads) % 4 || FF % 16 || F0 % 16 || cfg->compress != 2 || Q < 1 || cfg->n_sem < 1 ||
      cfg->n_sem > Q || cfg->kernel < 1 || cfg->kernel > 16)
    return mfail(SOPRO_ERR_INVALID, "unsupported Mimi encoder geometry (hidden=%d codebook_dim=%d)", C, Dc);
  MCK(cudaSetDevice(device));
  sopro_mimi_encoder* e = new sopro_mimi_encoder();
  e->device = device;
  e->cfg = *cfg;
  DevArena A;
  // conv weights [Cout][Cin][k] -> [Cout][(tap, ci)]; a strided conv's taps j = j2*r + rr are ordered (j2, rr, ci), which
  // is the same formula with k = 2r
  auto repack_conv = [&](const float* src, int cout, int cin, int k) {
    std::vector<float> r((size_t)cout * k * cin);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int j = 0; j < k; ++j) r[((size_t)co * k + j) * cin + ci] = src[((size_t)co * cin + ci) * k + j];
    return A.add(r.data(), r.size());
  };
  e->c0w = A.add(w->conv0_w, (size_t)F0 * cfg->kernel);
  e->c0b = A.add(w->conv0_b, F0);
  int ch = F0;
  for (int s = 0; s < cfg->n_ratios; ++s) {
    const sopro_mimi_enc_stage_weights_t& S = w->stage[s];
    sopro_mimi_encoder::Stage d;
    d.ratio = cfg->ratios[cfg->n_ratios - 1 - s];
    d.cin = ch;
    d.r1w = repack_conv(S.res1_w, ch / 2, ch, cfg->res_kernel);
    d.r1b = A.add(S.res1_b, ch / 2);
    d.r2w = repack_conv(S.res2_w, ch, ch / 2, 1);
    d.r2b = A.add(S.res2_b, ch);
    d.dw = repack_conv(S.down_w, 2 * ch, ch, 2 * d.ratio);
    d.db = A.add(S.down_b, 2 * ch);
    e->stages.push_back(d);
    ch *= 2;
  }
  e->lw = repack_conv(w->last_w, C, ch, cfg->last_kernel);
  e->lb = A.add(w->last_b, C);
  for (int l = 0; l < NL; ++l) {
    const sopro_mimi_layer_weights_t& L = w->layer[l];
    sopro_mimi::Layer d{};
    d.ln1w = A.add(L.ln1_w, C);
    d.ln1b = A.add(L.ln1_b, C);
    std::vector<float> qkv((size_t)3 * C * C);
    memcpy(qkv.data(), L.q_w, (size_t)C * C * 4);
    memcpy(qkv.data() + (size_t)C * C, L.k_w, (size_t)C * C * 4);
    memcpy(qkv.data() + (size_t)2 * C * C, L.v_w, (size_t)C * C * 4);
    d.qkv = A.add(qkv.data(), qkv.size());
    d.wo = A.add(L.o_w, (size_t)C * C);
    d.ls1 = A.add(L.ls1, C);
    d.ln2w = A.add(L.ln2_w, C);
    d.ln2b = A.add(L.ln2_b, C);
    d.fc1 = A.add(L.fc1_w, (size_t)FF * C);
    d.fc2 = A.add(L.fc2_w, (size_t)C * FF);
    d.ls2 = A.add(L.ls2, C);
    e->layers.push_back(d);
  }
  e->down_w = repack_conv(w->downsample_w, C, C, 4);
  {  // [2*Dc][C]: semantic input_proj rows, then acoustic
    std::vector<float> cat((size_t)2 * Dc * C);
    memcpy(cat.data(), w->sem_in_proj, (size_t)Dc * C * 4);
    memcpy(cat.data() + (size_t)Dc * C, w->ac_in_proj, (size_t)Dc * C * 4);
    e->inproj = A.add(cat.data(), cat.size());
  }
  e->embed = A.add(w->embed, (size_t)Q * V * Dc);
  cudaError_t err = cudaMalloc(&e->dev, A.host.size() * 4);
  if (err == cudaSuccess) err = cudaMemcpy(e->dev, A.host.data(), A.host.size() * 4, cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    if (e->dev) cudaFree(e->dev);
    delete e;
    return mfail(SOPRO_ERR_CUDA, "Mimi encoder weight upload failed: %s", cudaGetErrorString(err));
  }
  *out = e;
  return SOPRO_OK;
}

int sopro_mimi_encoder_destroy(sopro_mimi_encoder_t* e) {
  if (!e) return SOPRO_OK;
  cudaSetDevice(e->device);
  cudaFree(e->dev);
  cudaFree(e->rope);
  cudaFree(e->ws);
  cudaFree(e->wav_dev);
  cudaFree(e->codes_dev);
  cudaFree(e->lat_dev);
  delete e;
  return SOPRO_OK;
}

int64_t sopro_mimi_encoded_frames(const sopro_mimi_encoder_t* e, int64_t n_samples) {
  if (!e || n_samples < 1 || n_samples > kEncMaxSamples) return -1;
  return enc_plan(e->cfg, n_samples).T;
}

int sopro_mimi_encode(sopro_mimi_encoder_t* e, const float* wav, int64_t n_samples, int32_t* codes, float* latent, void* stream) {
  if (!e || !wav || !codes) return mfail(SOPRO_ERR_INVALID, "null argument");
  if (n_samples < 1 || n_samples > kEncMaxSamples)
    return mfail(SOPRO_ERR_INVALID, "n_samples=%lld outside [1, %lld]", (long long)n_samples, kEncMaxSamples);
  MCK(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const sopro_mimi_config_t& c = e->cfg;
  const int C = c.hidden, H = c.n_heads, Dh = C / H, FF = c.ffn, Dc = c.codebook_dim;
  const EncPlan P = enc_plan(c, n_samples);
  const int T2 = (int)P.len[c.n_ratios], T = (int)P.T;
  if (e->rope_T2 < T2) {  // RoPE table [cos rows | sin rows], as the decoder's
    cudaFree(e->rope);
    e->rope = nullptr;
    e->rope_T2 = 0;
    const int cap = std::max(T2, 256);
    std::vector<float> tab((size_t)2 * cap * (Dh / 2));
    for (int t = 0; t < cap; ++t)
      for (int d = 0; d < Dh / 2; ++d) {
        const float inv = 1.0f / powf(c.rope_theta, (float)(2 * d) / (float)Dh);
        const float f = (float)t * inv;
        tab[(size_t)t * (Dh / 2) + d] = cosf(f);
        tab[(size_t)(cap + t) * (Dh / 2) + d] = sinf(f);
      }
    MCK(cudaMalloc(&e->rope, tab.size() * 4));
    MCK(cudaMemcpyAsync(e->rope, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, st));
    MCK(cudaStreamSynchronize(st));
    e->rope_T2 = cap;
  }
  if (e->ws_bytes < P.need) {
    cudaFree(e->ws);
    e->ws = nullptr;
    e->ws_bytes = 0;
    cudaError_t ae = cudaMalloc(&e->ws, P.need);
    if (ae != cudaSuccess) return mfail(SOPRO_ERR_CUDA, "Mimi encoder workspace %zu MB: %s", P.need >> 20, cudaGetErrorString(ae));
    e->ws_bytes = P.need;
  }
  float* b0 = e->ws;
  float* b1 = b0 + P.buf;
  float* b2 = b1 + P.buf;
  float* x = b2 + P.buf;
  float* ln = x + P.xsz;
  const float* Wd = e->dev;
  GemmOp g{};
  // conv over rows [Tin][cin] (row stride cin), `taps` taps, left pad `pad` rows, M output rows
  auto conv = [&](const float* A, long long Tin, long long M, int cin, int taps, int pad, const float* W, const float* bias, int N,
                  float* Cc, int elu, int epi, const float* R, const float* scale) {
    g = GemmOp{};
    g.A = A; g.W = W; g.C = Cc; g.R = R; g.bias = bias; g.scale = scale;
    g.M = (int)M; g.N = N; g.K = taps * cin; g.Min = (int)Tin; g.Cin = cin; g.taps = taps; g.dil = 1; g.pad = pad; g.ldc = N;
    g.bias_mod = N; g.epi = epi; g.a_elu = elu;
    g.a_bs = Tin * cin; g.c_bs = M * N; g.r_bs = M * N;
    return launch_gemm(g, 1, st);
  };
  int rc;
  // ---- SEANet encoder
  float* cur = b0;   // stage input [len][ch]
  float* hid = b1;   // resblock hidden [len][ch/2]
  float* nxt = b2;
  {
    const long long tot = n_samples * c.num_filters;
    enc_conv0_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(wav, Wd + e->c0w, Wd + e->c0b, cur, n_samples, c.num_filters, c.kernel);
    MCK(cudaGetLastError());
  }
  int ch = c.num_filters;
  for (int s = 0; s < c.n_ratios; ++s) {
    const sopro_mimi_encoder::Stage& S = e->stages[s];
    const long long Ls = P.len[s], Lp = P.padded[s];
    const int r = S.ratio;
    // ResnetBlock (:412-451): x + conv1(ELU(conv3(ELU(x))))
    if ((rc = conv(cur, Ls, Ls, ch, c.res_kernel, c.res_kernel - 1, Wd + S.r1w, Wd + S.r1b, ch / 2, hid, 1, EPI_NONE, nullptr, nullptr))) return rc;
    if ((rc = conv(hid, Ls, Ls, ch / 2, 1, 0, Wd + S.r2w, Wd + S.r2b, ch, cur, 1, EPI_RES, cur, nullptr))) return rc;
    // ELU + conv kernel 2r stride r: zero rows up to a multiple of r, then 2 taps over [Lp/r][r*ch]
    if (Lp > Ls) MCK(cudaMemsetAsync(cur + (size_t)Ls * ch, 0, (size_t)(Lp - Ls) * ch * 4, st));
    if ((rc = conv(cur, Lp / r, Lp / r, r * ch, 2, 1, Wd + S.dw, Wd + S.db, 2 * ch, nxt, 1, EPI_NONE, nullptr, nullptr))) return rc;
    std::swap(cur, nxt);
    ch *= 2;
  }
  // ELU + conv k3 -> residual stream x [T2][C]
  if ((rc = conv(cur, T2, T2, ch, c.last_kernel, c.last_kernel - 1, Wd + e->lw, Wd + e->lb, C, x, 1, EPI_NONE, nullptr, nullptr))) return rc;
  // ---- encoder transformer (MimiTransformerLayer.forward :966-993), fp32 path of the decoder
  {
    const unsigned ln_grid = (unsigned)((T2 + 7) / 8);
    const size_t asm_bytes = (size_t)8 * (Dh + c.window) * 4;
    auto lin = [&](const float* A, int K, const float* W, int N, float* Cc, int epi, const float* R, const float* scale) {
      return conv(A, T2, T2, K, 1, 0, W, nullptr, N, Cc, 0, epi, R, scale);
    };
    for (const sopro_mimi::Layer& L : e->layers) {
      layernorm_kernel<<<ln_grid, 256, 0, st>>>(x, Wd + L.ln1w, Wd + L.ln1b, ln, (long long)T2, C, c.norm_eps);
      if ((rc = lin(ln, C, Wd + L.qkv, 3 * C, b0, EPI_NONE, nullptr, nullptr))) return rc;
      rope_kernel<<<dim3(T2, 1), 256, 0, st>>>(b0, e->rope, T2, e->rope_T2, C, H, 0, nullptr, nullptr, 1);
      attn_kernel<<<dim3((T2 + 7) / 8, H, 1), 256, asm_bytes, st>>>(b0, b1, T2, C, H, c.window, 0, nullptr, nullptr, 1);
      MCK(cudaGetLastError());
      if ((rc = lin(b1, C, Wd + L.wo, C, x, EPI_RES_SCALE, x, Wd + L.ls1))) return rc;
      layernorm_kernel<<<ln_grid, 256, 0, st>>>(x, Wd + L.ln2w, Wd + L.ln2b, ln, (long long)T2, C, c.norm_eps);
      if ((rc = lin(ln, C, Wd + L.fc1, FF, b0, EPI_GELU, nullptr, nullptr))) return rc;
      if ((rc = lin(b0, FF, Wd + L.fc2, C, x, EPI_RES_SCALE, x, Wd + L.ls2))) return rc;
    }
  }
  // ---- 25 -> 12.5 Hz: kernel 4, stride 2, no bias, replicate padding (2 rows left, 0 or 1 right)
  {
    const int rows = 2 * T + 2;
    const long long tot = (long long)rows * C;
    replicate_pad_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x, b0, T2, C, 2, rows);
    MCK(cudaGetLastError());
    float* lat = latent ? latent : b1;
    if ((rc = conv(b0, rows / 2, T, 2 * C, 2, 0, Wd + e->down_w, nullptr, C, lat, 0, EPI_NONE, nullptr, nullptr))) return rc;
    // ---- quantizer: both input projections in one GEMM, then the residual search
    if ((rc = conv(lat, T, T, C, 1, 0, Wd + e->inproj, nullptr, 2 * Dc, b2, 0, EPI_NONE, nullptr, nullptr))) return rc;
    rvq_encode_kernel<8><<<T, 256, 0, st>>>(b2, Wd + e->embed, codes, T, c.n_q, c.n_sem, c.vocab);
    MCK(cudaGetLastError());
  }
  return SOPRO_OK;
}

int sopro_mimi_encode_host(sopro_mimi_encoder_t* e, const float* wav_host, int64_t n_samples, int32_t* codes_host,
                           float* latent_host, void* stream) {
  if (!e || !wav_host || !codes_host) return mfail(SOPRO_ERR_INVALID, "null argument");
  if (n_samples < 1 || n_samples > kEncMaxSamples)
    return mfail(SOPRO_ERR_INVALID, "n_samples=%lld outside [1, %lld]", (long long)n_samples, kEncMaxSamples);
  MCK(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long T = enc_plan(e->cfg, n_samples).T;
  const size_t nc = (size_t)e->cfg.n_q * T, nl = (size_t)T * e->cfg.hidden;
  if (e->wav_cap < (size_t)n_samples) {
    cudaFree(e->wav_dev);
    e->wav_dev = nullptr;
    e->wav_cap = 0;
    MCK(cudaMalloc(&e->wav_dev, (size_t)n_samples * 4));
    e->wav_cap = (size_t)n_samples;
  }
  if (e->codes_cap < nc) {
    cudaFree(e->codes_dev);
    e->codes_dev = nullptr;
    e->codes_cap = 0;
    MCK(cudaMalloc(&e->codes_dev, nc * 4));
    e->codes_cap = nc;
  }
  if (latent_host && e->lat_cap < nl) {
    cudaFree(e->lat_dev);
    e->lat_dev = nullptr;
    e->lat_cap = 0;
    MCK(cudaMalloc(&e->lat_dev, nl * 4));
    e->lat_cap = nl;
  }
  MCK(cudaMemcpyAsync(e->wav_dev, wav_host, (size_t)n_samples * 4, cudaMemcpyHostToDevice, st));
  const int rc = sopro_mimi_encode(e, e->wav_dev, n_samples, e->codes_dev, latent_host ? e->lat_dev : nullptr, st);
  if (rc) return rc;
  MCK(cudaMemcpyAsync(codes_host, e->codes_dev, nc * 4, cudaMemcpyDeviceToHost, st));
  if (latent_host) MCK(cudaMemcpyAsync(latent_host, e->lat_dev, nl * 4, cudaMemcpyDeviceToHost, st));
  MCK(cudaStreamSynchronize(st));
  return SOPRO_OK;
}

}  // extern "C"
