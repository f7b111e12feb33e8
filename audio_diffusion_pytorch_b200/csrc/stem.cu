// Network-boundary and narrow-level kernels (CUDA cores; these levels are pure HBM streaming):
//   adp_stem_in     fp32 [B][C][T] (+append, +VDiffusion noising) -> k=s=f conv -> bf16 NWC
//   adp_stem_out    bf16 NWC -> nearest-up f + conv3 -> skip/gate -> v (+CFG, +sampler step, +loss)
//   adp_narrow_conv C == 8 ConvBlock: GN+SiLU -> conv3 (+residual) (+LayerNorm/FiLM) (+stats)
#include "common.cuh"
#include "ptx.cuh"

namespace adp {

constexpr int kStemMaxIn = 32;   // (cx+ca)*f
constexpr int kStemMaxC0 = 256;
constexpr int kStemMaxCo = 4;

// ------------------------------------------------------------------------------ stem_in
__global__ void __launch_bounds__(256) stem_in_kernel(const adp_stem_in_args a) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_w[];            // [c0][ci_total] then bias[c0]
  __shared__ float s_stats[2 * 64];
  const int cin = a.cx + a.ca;
  const int ci_total = cin * a.f;
  float* s_b = s_w + a.c0 * ci_total;
  for (int i = threadIdx.x; i < a.c0 * ci_total; i += blockDim.x) s_w[i] = a.w[i];
  for (int i = threadIdx.x; i < a.c0; i += blockDim.x) s_b[i] = a.bias ? a.bias[i] : 0.f;
  if (threadIdx.x < 128) s_stats[threadIdx.x] = 0.f;
  __syncthreads();

  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int To = a.T / a.f;
  const int to = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = to < To;
  float in[kStemMaxIn];
  float al = 1.f, be = 0.f;
  if (a.noise) { al = a.alpha[b]; be = a.beta[b]; }
#pragma unroll
  for (int i = 0; i < kStemMaxIn; ++i) in[i] = 0.f;
  if (ok) {
    // PyTorch Conv1d weight layout [c0][cin][f]: input index i = c*f + j
#pragma unroll
    for (int i = 0; i < kStemMaxIn; ++i) {
      if (i < ci_total) {
        const int c = i / a.f, j = i - c * a.f;
        const size_t tt = static_cast<size_t>(to) * a.f + j;
        if (c < a.cx) {
          const size_t idx = (static_cast<size_t>(b) * a.cx + c) * a.T + tt;
          float v = a.x[idx];
          if (a.noise) v = al * v + be * a.noise[idx];   // reference diffusion.py:91
          in[i] = v;
        } else {
          in[i] = a.append[(static_cast<size_t>(b) * a.ca + (c - a.cx)) * a.T + tt];
        }
      }
    }
  }
  const int gsz = a.stats ? a.c0 / a.groups : 1;
  GroupStatAcc acc;
  __nv_bfloat16* orow =
      static_cast<__nv_bfloat16*>(a.out) + (static_cast<size_t>(b) * To + (ok ? to : 0)) * a.c0;
  for (int co = 0; co < a.c0; co += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = s_b[co + j];
#pragma unroll
    for (int i = 0; i < kStemMaxIn; ++i) {
      if (i < ci_total) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += in[i] * s_w[(co + j) * ci_total + i];
      }
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
    o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    if (ok) *reinterpret_cast<uint4*>(orow + co) = o;
    if (a.stats) {
      const uint32_t ou[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 r = unpack_bf16(ou[j]);
        acc.add(ok ? r.x : 0.f, (co + 2 * j) / gsz, s_stats, lane);
        acc.add(ok ? r.y : 0.f, (co + 2 * j + 1) / gsz, s_stats, lane);
      }
    }
  }
  if (a.stats) {
    acc.flush(s_stats, lane);
    __syncthreads();
    if (threadIdx.x < 2 * a.groups && s_stats[threadIdx.x] != 0.f)
      atomicAdd(a.stats + static_cast<size_t>(b) * 2 * a.groups + threadIdx.x,
                static_cast<double>(s_stats[threadIdx.x]));
  }
}

// ----------------------------------------------------------------------------- stem_out
// conv3 on the nearest-upsampled h for one output position; w in smem as [co][k][c0]
__device__ __forceinline__ void stem_out_conv(const __nv_bfloat16* __restrict__ hb, int T, int f,
                                              int c0, int co_n, int t, const float* s_w,
                                              float (&y)[kStemMaxCo]) {
  const int Tl = T / f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int idx = t + k - 1;
    if (idx < 0 || idx >= T) continue;       // zero padding of the upsampled signal
    const int q = idx / f;                   // nearest-neighbour source row
    (void)Tl;
    const uint4* row = reinterpret_cast<const uint4*>(hb + static_cast<size_t>(q) * c0);
    for (int c8 = 0; c8 < c0; c8 += 8) {
      const uint4 u = __ldg(row + (c8 >> 3));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
      float hv[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f2 = unpack_bf16(uu[j]);
        hv[2 * j] = f2.x; hv[2 * j + 1] = f2.y;
      }
#pragma unroll
      for (int o = 0; o < kStemMaxCo; ++o) {
        if (o < co_n) {
          const float* wp = s_w + (o * 3 + k) * c0 + c8;
#pragma unroll
          for (int j = 0; j < 8; ++j) y[o] += hv[j] * wp[j];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) stem_out_kernel(const adp_stem_out_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const int ldg = a.ld_gate > 0 ? a.ld_gate : a.co;
  extern __shared__ float s_w[];   // conv w [co][3][c0], bias[co], adapt w [co][cin], adapt b[co]
  __shared__ double s_loss[8];
  const int cin = a.cx + a.ca;
  float* s_b = s_w + a.co * 3 * a.c0;
  float* s_wa = s_b + a.co;
  float* s_ba = s_wa + a.co * cin;
  for (int i = threadIdx.x; i < a.co * 3 * a.c0; i += blockDim.x) {
    // PyTorch layout [co][c0][3] -> [co][3][c0]
    const int o = i / (3 * a.c0), r = i - o * 3 * a.c0, k = r / a.c0, c = r - k * a.c0;
    s_w[i] = a.w[(o * a.c0 + c) * 3 + k];
  }
  for (int i = threadIdx.x; i < a.co; i += blockDim.x) {
    s_b[i] = a.bias ? a.bias[i] : 0.f;
    s_ba[i] = a.b_adapt ? a.b_adapt[i] : 0.f;
  }
  if (a.w_adapt)
    for (int i = threadIdx.x; i < a.co * cin; i += blockDim.x) s_wa[i] = a.w_adapt[i];
  __syncthreads();

  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int Tl = a.T / a.f;
  double lsum = 0.0;
  if (t < a.T) {
    float al = 1.f, be = 0.f;
    if (a.noise) { al = a.alpha[b]; be = a.beta[b]; }
    // block input at this position (the U-Net skip): cat([x(_noisy), append])
    float xin[8];
    float xraw[kStemMaxCo], nraw[kStemMaxCo];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      xin[c] = 0.f;
      if (c < a.cx) {
        const size_t idx = (static_cast<size_t>(b) * a.cx + c) * a.T + t;
        float v = a.x[idx];
        if (c < kStemMaxCo) { xraw[c] = v; nraw[c] = 0.f; }
        if (a.noise) {
          const float nv = a.noise[idx];
          if (c < kStemMaxCo) nraw[c] = nv;
          v = al * v + be * nv;
        }
        xin[c] = v;
      } else if (c < cin) {
        xin[c] = a.append[(static_cast<size_t>(b) * a.ca + (c - a.cx)) * a.T + t];
      }
    }
    float y[kStemMaxCo], ym[kStemMaxCo];
#pragma unroll
    for (int o = 0; o < kStemMaxCo; ++o) { y[o] = o < a.co ? s_b[o] : 0.f; ym[o] = y[o]; }
    const __nv_bfloat16* h = static_cast<const __nv_bfloat16*>(a.h);
    stem_out_conv(h + static_cast<size_t>(b) * Tl * a.c0, a.T, a.f, a.c0, a.co, t, s_w, y);
    if (a.cfg)
      stem_out_conv(h + static_cast<size_t>(b + a.B) * Tl * a.c0, a.T, a.f, a.c0, a.co, t, s_w, ym);
#pragma unroll
    for (int o = 0; o < kStemMaxCo; ++o) {
      if (o < a.co) {
        float skip;
        if (a.w_adapt) {                       // SkipAdapter 1x1 conv (in != out channels)
          skip = s_ba[o];
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < cin) skip += xin[c] * s_wa[o * cin + c];
        } else {
          skip = xin[o];
        }
        float v = skip + a.gate[static_cast<size_t>(b) * ldg + o] * y[o];   // MergeModulate
        if (a.cfg) {
          const float vm = skip + a.gate[static_cast<size_t>(b + a.B) * ldg + o] * ym[o];
          v = vm + (v - vm) * a.cfg_scale;                                   // CFG combine
        }
        const size_t oidx = (static_cast<size_t>(b) * a.co + o) * a.T + t;
        if (a.v_out) a.v_out[oidx] = v;
        if (a.x_next) {                       // reference diffusion.py:185-187
          const float a0 = a.ab[0], b0 = a.ab[1], a1 = a.ab[2], b1 = a.ab[3];
          const float xv = xin[o];
          const float x_pred = a0 * xv - b0 * v;
          const float n_pred = b0 * xv + a0 * v;
          a.x_next[oidx] = a1 * x_pred + b1 * n_pred;
        }
        if (a.loss_sum) {                     // reference diffusion.py:92,95
          const float vt = al * nraw[o] - be * xraw[o];
          const float d = v - vt;
          lsum += static_cast<double>(d) * d;
          if (a.dv) a.dv[oidx] = 2.f * d / (static_cast<float>(a.B) * a.co * a.T);
        }
      }
    }
  }
  if (a.loss_sum) {
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if ((threadIdx.x & 31) == 0) s_loss[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < (blockDim.x >> 5); ++i) tot += s_loss[i];
      atomicAdd(a.loss_sum, tot);
    }
  }
}

// -------------------------------------------------------------------------- narrow_conv
template <int C>
__global__ void __launch_bounds__(256) narrow_conv_kernel(const adp_narrow_conv_args a) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int TB = 256;
  __shared__ __align__(16) float s_in[(TB + 2) * C];
  __shared__ __align__(16) float s_w[3 * C * C];   // [k][ci][co]
  __shared__ float s_a[C], s_d[C], s_b[C];
  __shared__ float s_stats[2 * 64];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TB;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 128) s_stats[threadIdx.x] = 0.f;
  for (int i = threadIdx.x; i < 3 * C * C; i += TB) {
    const int k = i / (C * C), r = i - k * C * C, ci = r / C, co = r - ci * C;
    s_w[i] = a.w[(co * C + ci) * 3 + k];           // PyTorch [co][ci][k]
  }
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    const int gsz = C / a.groups, g = c / gsz;
    const double inv_n = 1.0 / (static_cast<double>(gsz) * a.T);
    const double s = a.stats_in[(static_cast<size_t>(b) * a.groups + g) * 2];
    const double q = a.stats_in[(static_cast<size_t>(b) * a.groups + g) * 2 + 1];
    const double mean = s * inv_n;
    double var = q * inv_n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.gn_eps)));
    const float ga = a.gamma[c] * rstd;
    s_a[c] = ga;
    s_d[c] = a.beta[c] - static_cast<float>(mean) * ga;
    s_b[c] = a.bias ? a.bias[c] : 0.f;
  }
  __syncthreads();

  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(a.x) + static_cast<size_t>(b) * a.T * C;
  // activated input rows t0-1 .. t0+TB into smem (zero outside [0,T): conv padding)
  for (int i = threadIdx.x; i < TB + 2; i += TB) {
    const int t = t0 - 1 + i;
    float v[C];
    if (t >= 0 && t < a.T) {
#pragma unroll
      for (int c8 = 0; c8 < C; c8 += 8) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(t) * C + c8));
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f2 = unpack_bf16(uu[j]);
          v[c8 + 2 * j] = silu_f(f2.x * s_a[c8 + 2 * j] + s_d[c8 + 2 * j]);
          v[c8 + 2 * j + 1] = silu_f(f2.y * s_a[c8 + 2 * j + 1] + s_d[c8 + 2 * j + 1]);
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < C; c += 4)
      *reinterpret_cast<float4*>(&s_in[i * C + c]) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
  }
  __syncthreads();

  const int t = t0 + threadIdx.x;
  const bool ok = t < a.T;
  float y[C];
#pragma unroll
  for (int c = 0; c < C; ++c) y[c] = s_b[c];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int ci = 0; ci < C; ++ci) {
      const float xv = s_in[(threadIdx.x + k) * C + ci];
#pragma unroll
      for (int co = 0; co < C; co += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(&s_w[(k * C + ci) * C + co]);
        y[co] += xv * w4.x; y[co + 1] += xv * w4.y; y[co + 2] += xv * w4.z; y[co + 3] += xv * w4.w;
      }
    }
  }
  const size_t roff = (static_cast<size_t>(b) * a.T + (ok ? t : 0)) * C;
  if (a.residual && ok) {
#pragma unroll
    for (int c8 = 0; c8 < C; c8 += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(
          static_cast<const __nv_bfloat16*>(a.residual) + roff + c8);
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f2 = unpack_bf16(uu[j]);
        y[c8 + 2 * j] += f2.x; y[c8 + 2 * j + 1] += f2.y;
      }
    }
  }
  if (a.scale_shift) {   // following ModulationItem: LayerNorm over C (no affine) + FiLM
    float mean = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) mean += y[c];
    mean *= (1.f / C);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { const float d = y[c] - mean; var += d * d; }
    const float rstd = rsqrtf(var * (1.f / C) + a.ln_eps);
    const float* ss = a.scale_shift + static_cast<size_t>(b) * a.ss_stride;
#pragma unroll
    for (int c = 0; c < C; ++c) y[c] = (y[c] - mean) * rstd * (1.f + ss[c]) + ss[C + c];
  }
  const int gszo = a.stats_out ? C / a.groups : 1;
  GroupStatAcc acc;
#pragma unroll
  for (int c8 = 0; c8 < C; c8 += 8) {
    uint4 o;
    o.x = pack_bf16(y[c8], y[c8 + 1]); o.y = pack_bf16(y[c8 + 2], y[c8 + 3]);
    o.z = pack_bf16(y[c8 + 4], y[c8 + 5]); o.w = pack_bf16(y[c8 + 6], y[c8 + 7]);
    if (ok) *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.y) + roff + c8) = o;
    if (a.stats_out) {
      const uint32_t ou[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 r = unpack_bf16(ou[j]);
        acc.add(ok ? r.x : 0.f, (c8 + 2 * j) / gszo, s_stats, lane);
        acc.add(ok ? r.y : 0.f, (c8 + 2 * j + 1) / gszo, s_stats, lane);
      }
    }
  }
  if (a.stats_out) {
    acc.flush(s_stats, lane);
    __syncthreads();
    if (threadIdx.x < 2 * a.groups && s_stats[threadIdx.x] != 0.f)
      atomicAdd(a.stats_out + static_cast<size_t>(b) * 2 * a.groups + threadIdx.x,
                static_cast<double>(s_stats[threadIdx.x]));
  }
}

}  // namespace adp

using namespace adp;

extern "C" int adp_stem_in(const adp_stem_in_args* args, adp_stream_t stream) {
  ADP_CHECK(args && args->x && args->w && args->out, "adp_stem_in: null pointer");
  const adp_stem_in_args& a = *args;
  ADP_CHECK((a.cx + a.ca) * a.f <= kStemMaxIn && a.f >= 1 && a.T % a.f == 0,
            "adp_stem_in: (cx+ca)*f = %d > %d or T %% f != 0", (a.cx + a.ca) * a.f, kStemMaxIn);
  ADP_CHECK(a.c0 % 8 == 0 && a.c0 <= kStemMaxC0, "adp_stem_in: c0=%d unsupported", a.c0);
  ADP_CHECK((a.ca == 0) == (a.append == nullptr), "adp_stem_in: append / ca mismatch");
  ADP_CHECK(!a.noise || (a.alpha && a.beta), "adp_stem_in: noise needs alpha/beta");
  if (a.stats) ADP_CHECK(a.groups > 0 && a.groups <= 64 && a.c0 % a.groups == 0, "adp_stem_in: groups");
  const size_t smem = (static_cast<size_t>(a.c0) * (a.cx + a.ca) * a.f + a.c0) * sizeof(float);
  dim3 grid((a.T / a.f + 255) / 256, a.B);
  ADP_CUDA(launch_k(stem_in_kernel, grid, dim3(256), smem, as_stream(stream), a));
  ADP_LAUNCH_CHECK();
  return 0;
}

extern "C" int adp_stem_out(const adp_stem_out_args* args, adp_stream_t stream) {
  ADP_CHECK(args && args->h && args->x && args->w && args->gate, "adp_stem_out: null pointer");
  const adp_stem_out_args& a = *args;
  ADP_CHECK(a.co >= 1 && a.co <= kStemMaxCo && a.cx + a.ca <= 8 && a.co <= a.cx,
            "adp_stem_out: co=%d cx=%d ca=%d unsupported", a.co, a.cx, a.ca);
  ADP_CHECK(a.c0 % 8 == 0 && a.c0 <= kStemMaxC0 && a.f >= 1 && a.T % a.f == 0,
            "adp_stem_out: c0=%d f=%d", a.c0, a.f);
  ADP_CHECK(a.w_adapt || a.cx + a.ca == a.co, "adp_stem_out: identity skip needs cx+ca == co");
  ADP_CHECK((a.ca == 0) == (a.append == nullptr), "adp_stem_out: append / ca mismatch");
  ADP_CHECK(!a.x_next || a.ab, "adp_stem_out: x_next needs ab");
  ADP_CHECK(!a.loss_sum || (a.noise && a.alpha && a.beta), "adp_stem_out: loss needs noise/alpha/beta");
  const size_t smem =
      (static_cast<size_t>(a.co) * 3 * a.c0 + 2 * a.co + a.co * (a.cx + a.ca)) * sizeof(float);
  dim3 grid((a.T + 255) / 256, a.B);
  ADP_CUDA(launch_k(stem_out_kernel, grid, dim3(256), smem, as_stream(stream), a));
  ADP_LAUNCH_CHECK();
  return 0;
}

extern "C" int adp_narrow_conv(const adp_narrow_conv_args* args, adp_stream_t stream) {
  ADP_CHECK(args && args->x && args->y && args->stats_in && args->gamma && args->beta && args->w,
            "adp_narrow_conv: null pointer");
  const adp_narrow_conv_args& a = *args;
  ADP_CHECK(a.C == 8, "adp_narrow_conv: only C == 8 is built (C=%d); wider levels use adp_conv_gemm",
            a.C);
  ADP_CHECK(a.groups > 0 && a.C % a.groups == 0, "adp_narrow_conv: groups=%d", a.groups);
  dim3 grid((a.T + 255) / 256, a.B);
  ADP_CUDA(launch_k(narrow_conv_kernel<8>, grid, dim3(256), (size_t)0, as_stream(stream), a));
  ADP_LAUNCH_CHECK();
  return 0;
}
