// v1 (non-persistent) shifted-tap GEMM, kept as the A/B reference for conv_gemm.cu (adp_debug_set(0, 1)).
// adp_conv_gemm: shifted-tap GEMM on tcgen05 (see include/adp_b200.h).
//
// GEMM view (swapped w.r.t. the usual conv-as-GEMM so that one accumulator ROW is one time
// position): D[m = time position, n = output channel] = sum_{tap,k} A[m + off(tap), k] * W[n, tap, k]
//   A: channels-last activations  -> K-major operand, TMA box [BK x 128 rows], rows shifted
//      per tap; rows outside [0,T) are zero-filled by TMA = the conv's zero padding.
//   W: packed weights [N][taps*C_in] -> K-major operand, TMA box [BK x BN].
//   D: fp32 in TMEM, 128 lanes (rows) x BN columns.
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 TMA
// producer + TMEM allocator, warp 5 MMA issuer.  smem ring of `stages` {A,W} tiles with
// full/empty mbarriers; tcgen05.commit releases ring slots and publishes the accumulator.
#include "common.cuh"
#include "ptx.cuh"

namespace adp {

constexpr int kBMv1 = 128;
constexpr int kMaxStagesV1 = 6;
constexpr int kMaxGroupsV1 = 32;

struct GemmV1Params {
  __nv_bfloat16* out;
  const __nv_bfloat16* residual;
  const float* bias;
  const float* gate;
  double* stats;
  int T, tiles_per_batch, c_in, ldo;
  int n_pad, n_valid;
  int ntaps, tap_off0, tap_off1, tap_off2, up_factor;
  int groups, group_size;
  int stages;
  int out_fp32;
  int ld_gate;
};

template <int BN, int SW>
struct GemmV1Cfg {
  static constexpr int BK = SW / 2;             // bf16 elements per swizzle row
  static constexpr int A_BYTES = kBMv1 * SW;
  static constexpr int W_TX_BYTES = BN * SW;    // bytes TMA actually writes
  static constexpr int W_BYTES = (W_TX_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int CH = BN < 32 ? 16 : 32;  // epilogue column chunk
};

template <int BN, int SW>
__global__ void __launch_bounds__(192)
conv_gemm_v1_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const GemmV1Params p) {
  using Cfg = GemmV1Cfg<BN, SW>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStagesV1];
  __shared__ uint64_t empty_bar[kMaxStagesV1];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_stats[2 * kMaxGroupsV1];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* tiles = smem_raw + ((1024u - (raw & 1023u)) & 1023u);

  const int m_tile = blockIdx.x;
  const int b = m_tile / p.tiles_per_batch;
  const int t0 = (m_tile - b * p.tiles_per_batch) * kBMv1;
  const int n0 = blockIdx.y * BN;          // in the padded [phases*n_pad] space
  const int phase = n0 / p.n_pad;
  const int ch0 = n0 - phase * p.n_pad;    // channel offset inside the phase

  int ntaps = p.ntaps, off0 = p.tap_off0, off1 = p.tap_off1, off2 = p.tap_off2;
  if (p.up_factor > 1) {  // nearest-upsample + conv3: taps collapse per output phase
    if (phase == 0) { ntaps = 2; off0 = -1; off1 = 0; }
    else if (phase == p.up_factor - 1) { ntaps = 2; off0 = 0; off1 = 1; }
    else { ntaps = 1; off0 = 0; }
  }
  const int k_chunks = p.c_in / Cfg::BK;
  const int iters = ntaps * k_chunks;
  const int stages = p.stages;

  if (threadIdx.x < 2 * kMaxGroupsV1) s_stats[threadIdx.x] = 0.f;
  if (warp == 4) {
    tmem_alloc(&tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  } else if (warp == 5 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int tap = it / k_chunks;
        const int kc = it - tap * k_chunks;
        const int off = tap == 0 ? off0 : (tap == 1 ? off1 : off2);
        uint8_t* a_s = tiles + s * Cfg::STAGE_BYTES;
        uint8_t* w_s = a_s + Cfg::A_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], Cfg::A_BYTES + Cfg::W_TX_BYTES);
        tma_load_3d(a_s, &tmA, &full_bar[s], kc * Cfg::BK, t0 + off, b);
        tma_load_2d(w_s, &tmW, &full_bar[s], tap * p.c_in + kc * Cfg::BK, n0);
      }
    }
  } else if (warp == 5) {
    // -------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBMv1, BN, 0, 0);
      for (int it = 0; it < iters; ++it) {
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + s * Cfg::STAGE_BYTES);
        const uint32_t w_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int kk = 0; kk < Cfg::BK / 16; ++kk) {
          umma_bf16(tmem_base, umma_desc_kmajor<SW>(a_addr + kk * 32),
                    umma_desc_kmajor<SW>(w_addr + kk * 32), idesc, (it | kk) != 0);
        }
        umma_commit(&empty_bar[s]);  // slot reusable once these MMAs have read it
      }
      umma_commit(&tmem_full_bar);   // accumulator complete
    }
  } else {
    // ---------------------------------------------------------------------- epilogue
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const int t = t0 + row;
    const bool row_ok = t < p.T;
    const size_t row_off = (static_cast<size_t>(b) * p.T + (row_ok ? t : 0)) * p.ldo +
                           static_cast<size_t>(phase) * p.n_valid;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const bool do_stats = p.stats != nullptr;
    GroupStatAcc acc;

#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += Cfg::CH) {
      uint32_t r[Cfg::CH];
      if constexpr (Cfg::CH == 16) tmem_ld16(taddr + c0, r);
      else tmem_ld32(taddr + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int v8 = 0; v8 < Cfg::CH / 8; ++v8) {
        const int ch = ch0 + c0 + v8 * 8;       // first channel of this 8-vector
        if (ch >= p.n_valid) continue;           // padded columns (uniform branch)
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[v8 * 8 + j]);
        if (p.bias) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ch));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + ch + 4));
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
          v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        if (p.gate) {
          const float* gp = p.gate + static_cast<size_t>(b) * p.ld_gate + ch;
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(gp + 4));
          v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w;
          v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
        }
        if (p.out_fp32) {
          if (row_ok) {
            float* op = reinterpret_cast<float*>(p.out) + row_off + ch;
            *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          continue;
        }
        if (row_ok) {
          if (p.residual) {
            const uint4 rr = *reinterpret_cast<const uint4*>(p.residual + row_off + ch);
            const float2 r0 = unpack_bf16(rr.x), r1 = unpack_bf16(rr.y);
            const float2 r2 = unpack_bf16(rr.z), r3 = unpack_bf16(rr.w);
            v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y;
            v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
          }
          uint4 o;
          o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
          o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
          *reinterpret_cast<uint4*>(p.out + row_off + ch) = o;
          if (do_stats) {  // statistics of the ROUNDED values the next layer will read
            const float2 q0 = unpack_bf16(o.x), q1 = unpack_bf16(o.y);
            const float2 q2 = unpack_bf16(o.z), q3 = unpack_bf16(o.w);
            v[0] = q0.x; v[1] = q0.y; v[2] = q1.x; v[3] = q1.y;
            v[4] = q2.x; v[5] = q2.y; v[6] = q3.x; v[7] = q3.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;
        }
        if (do_stats) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc.add(v[j], (ch + j) / p.group_size, s_stats, lane);
        }
      }
    }
    if (do_stats) {
      acc.flush(s_stats, lane);
      named_bar_sync(1, 128);
      if (threadIdx.x < 2 * p.groups) {
        const float val = s_stats[threadIdx.x];
        if (val != 0.f)
          atomicAdd(p.stats + static_cast<size_t>(b) * 2 * p.groups + threadIdx.x,
                    static_cast<double>(val));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int SW>
static int launch_gemm_v1(const adp_conv_gemm_args& a, cudaStream_t stream) {
  using Cfg = GemmV1Cfg<BN, SW>;
  const int tiles_per_batch = (a.T + kBMv1 - 1) / kBMv1;
  const int max_taps = a.up_factor > 1 ? 2 : a.ntaps;
  const int iters_max = max_taps * (a.c_in / Cfg::BK);

  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[3] = {(uint64_t)a.c_in, (uint64_t)a.T, (uint64_t)a.B};
    const uint64_t strides[2] = {(uint64_t)a.lda * 2, (uint64_t)a.T * a.lda * 2};
    const uint32_t box[3] = {(uint32_t)Cfg::BK, (uint32_t)kBMv1, 1};
    if (int e = make_tmap_bf16(&tmA, a.a, 3, dims, strides, box, SW)) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.k_total, (uint64_t)a.phases * a.n_pad};
    const uint64_t strides[1] = {(uint64_t)a.k_total * 2};
    const uint32_t box[2] = {(uint32_t)Cfg::BK, (uint32_t)BN};
    if (int e = make_tmap_bf16(&tmW, a.w, 2, dims, strides, box, SW)) return e;
  }

  int stages = iters_max < 4 ? iters_max : 4;
  // keep several CTAs per SM resident on the short-K (bandwidth-bound) shapes
  while (stages > 2 && stages * Cfg::STAGE_BYTES > 96 * 1024 && iters_max <= 6) --stages;
  if (stages * Cfg::STAGE_BYTES > 200 * 1024) stages = (200 * 1024) / Cfg::STAGE_BYTES;
  if (stages < 1) stages = 1;
  const size_t smem = (size_t)stages * Cfg::STAGE_BYTES + 1024;

  static size_t smem_attr = 0;   // opt-in dynamic smem (static smem counts against the 227 KB)
  if (smem > smem_attr) {
    ADP_CUDA(cudaFuncSetAttribute(conv_gemm_v1_kernel<BN, SW>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_attr = smem;
  }

  GemmV1Params p;
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.residual = static_cast<const __nv_bfloat16*>(a.residual);
  p.bias = a.bias;
  p.gate = a.gate;
  p.stats = a.stats;
  p.T = a.T;
  p.tiles_per_batch = tiles_per_batch;
  p.c_in = a.c_in;
  p.ldo = a.ldo;
  p.n_pad = a.n_pad;
  p.n_valid = a.n_valid;
  p.ntaps = a.ntaps;
  p.tap_off0 = a.tap_off[0];
  p.tap_off1 = a.tap_off[1];
  p.tap_off2 = a.tap_off[2];
  p.up_factor = a.up_factor;
  p.groups = a.stats ? a.groups : 0;
  p.group_size = a.stats ? a.n_valid / a.groups : 1;
  p.stages = stages;
  p.out_fp32 = a.out_fp32;
  p.ld_gate = a.ld_gate > 0 ? a.ld_gate : a.n_valid;

  dim3 grid(a.B * tiles_per_batch, a.phases * a.n_pad / BN);
  conv_gemm_v1_kernel<BN, SW><<<grid, 192, smem, stream>>>(tmA, tmW, p);
  ADP_LAUNCH_CHECK();
  return 0;
}

template <int SW>
static int dispatch_bn_v1(const adp_conv_gemm_args& a, int bn, cudaStream_t s) {
  switch (bn) {
    case 16: return launch_gemm_v1<16, SW>(a, s);
    case 32: return launch_gemm_v1<32, SW>(a, s);
    case 64: return launch_gemm_v1<64, SW>(a, s);
    case 128: return launch_gemm_v1<128, SW>(a, s);
    case 256: return launch_gemm_v1<256, SW>(a, s);
  }
  return set_error("adp_conv_gemm: unsupported N tile %d", bn);
}

}  // namespace adp

namespace adp {
int conv_gemm_v1(const adp_conv_gemm_args* args, adp_stream_t stream) {
  ADP_CHECK(args != nullptr, "adp_conv_gemm: null args");
  const adp_conv_gemm_args& a = *args;
  ADP_CHECK(a.a && a.w && a.out, "adp_conv_gemm: null a/w/out");
  ADP_CHECK(a.B > 0 && a.T > 0, "adp_conv_gemm: bad B=%d T=%d", a.B, a.T);
  ADP_CHECK(a.c_in >= 16 && a.c_in % 16 == 0, "adp_conv_gemm: c_in=%d must be a multiple of 16",
            a.c_in);
  ADP_CHECK(a.lda % 8 == 0 && a.ldo % 8 == 0 && a.k_total % 8 == 0 && a.lda >= a.c_in,
            "adp_conv_gemm: pitches lda=%d ldo=%d k_total=%d must be multiples of 8", a.lda, a.ldo,
            a.k_total);
  ADP_CHECK(a.n_valid > 0 && a.n_valid % 8 == 0 && a.n_valid <= a.n_pad && a.n_pad % 16 == 0,
            "adp_conv_gemm: n_valid=%d n_pad=%d", a.n_valid, a.n_pad);
  ADP_CHECK(a.phases >= 1, "adp_conv_gemm: phases=%d", a.phases);
  if (a.up_factor > 1) {
    ADP_CHECK(a.phases == a.up_factor, "adp_conv_gemm: phases (%d) != up_factor (%d)", a.phases,
              a.up_factor);
    ADP_CHECK(a.k_total >= 2 * a.c_in, "adp_conv_gemm: upsample weights need 2 tap slots");
  } else {
    ADP_CHECK(a.ntaps >= 1 && a.ntaps <= 3 && a.phases == 1, "adp_conv_gemm: ntaps=%d phases=%d",
              a.ntaps, a.phases);
    ADP_CHECK(a.k_total >= a.ntaps * a.c_in, "adp_conv_gemm: k_total too small");
  }
  ADP_CHECK(a.ld_gate % 4 == 0, "adp_conv_gemm: ld_gate=%d must be a multiple of 4", a.ld_gate);
  if (a.out_fp32) {
    ADP_CHECK(!a.residual && !a.stats, "adp_conv_gemm: out_fp32 excludes residual/stats");
  }
  if (a.stats) {
    ADP_CHECK(a.groups > 0 && a.groups <= kMaxGroupsV1 && a.n_valid % a.groups == 0,
              "adp_conv_gemm: groups=%d n_valid=%d", a.groups, a.n_valid);
  }
  // N tile: largest that divides n_pad and still gives >= ~1 wave of CTAs
  int bn = a.block_n;
  if (bn == 0) {
    const long m_tiles = (long)a.B * ((a.T + kBMv1 - 1) / kBMv1);
    bn = 16;
    for (int cand = 256; cand >= 16; cand >>= 1) {
      if (a.n_pad % cand) continue;
      const long ctas = m_tiles * (a.phases * a.n_pad / cand);
      if (ctas >= 120 || cand <= 64) { bn = cand; break; }
    }
  }
  ADP_CHECK(a.n_pad % bn == 0, "adp_conv_gemm: N tile %d does not divide n_pad %d", bn, a.n_pad);
  cudaStream_t s = as_stream(stream);
  if (a.c_in % 64 == 0) return dispatch_bn_v1<128>(a, bn, s);
  if (a.c_in % 32 == 0) return dispatch_bn_v1<64>(a, bn, s);
  return dispatch_bn_v1<32>(a, bn, s);
}
}  // namespace adp
