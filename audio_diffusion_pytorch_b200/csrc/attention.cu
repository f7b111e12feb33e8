// adp_attention: softmax(q k^T * scale) v, head dim 64, on tcgen05 (a_unet AttentionBase).
//
// One CTA = one (batch, head, 128-query tile); two CTAs are resident per SM so one CTA's
// softmax overlaps the other's MMAs.  S = Q K^T (128 x 128 keys) and O (128 x 64) accumulate
// in TMEM.  Each query row is owned by TWO softmax threads (same TMEM lane, one per 64-key
// half), so the online softmax needs no shuffles: the halves exchange their row max through
// shared memory once per key tile and their row sums once at the end.
//   warps 0-7: softmax     S(TMEM, read once) -> max -> ex2 -> P(bf16, swizzled smem);
//                          O(TMEM) rescaled only when a row max actually moved
//   warp 8   : TMA producer  Q once; K double-buffered; V single-buffered (it is needed last)
//   warp 9   : MMA issuer    S = Q K_j^T ; (wait P) ; O += P_j V_j
// K is the K-major B operand of the first GEMM (keys x d); V is used in place as the
// MN-major B operand of the second (d contiguous per key), so no transpose is materialised.
#include "common.cuh"
#include "ptx.cuh"

namespace adp {

constexpr int kQT = 128;       // query rows per CTA
constexpr int kKT = 128;       // keys per tile
constexpr int kD = 64;         // head dim
constexpr int kQBytes = kQT * kD * 2;          // 16 KB
constexpr int kKVBytes = kKT * kD * 2;         // 16 KB
constexpr int kPBytes = kQT * kKT * 2;         // 32 KB (two 64-key swizzle chunks)
constexpr int kAttnSmem = kQBytes + 2 * kKVBytes + kKVBytes + kPBytes + 1024;   // 97 KB
constexpr uint32_t kTmemCols = 256;            // S: [0,128)  O: [128,192)
constexpr int kSoftmaxThreads = 256;

struct AttnParams {
  __nv_bfloat16* o;
  float* lse;         // optional [B][H][Tq]: log-sum-exp of the scaled scores (training forward)
  int Tq, Tk, ldo;
  float scale_log2;   // scale * log2(e)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(320, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, s_full, p_full, o_done, v_full, v_empty;
  __shared__ uint64_t k_full[2], k_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_xch[2][kQT];

  pdl_launch_dependents();
  const int warp = warp_id_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* q_s = base;
  uint8_t* k_s = q_s + kQBytes;            // [2][16 KB]
  uint8_t* v_s = k_s + 2 * kKVBytes;       // 16 KB
  uint8_t* p_s = v_s + kKVBytes;           // 32 KB

  const int t0 = blockIdx.x * kQT;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (p.Tk + kKT - 1) / kKT;

  if (warp == 8) {
    tmem_alloc(&tmem_slot, kTmemCols);
    tmem_relinquish();
  } else if (warp == 9 && lane == 0) {
    mbar_init(&q_full, 1);
    mbar_init(&s_full, 1);
    mbar_init(&p_full, kSoftmaxThreads);
    mbar_init(&o_done, 1);
    mbar_init(&v_full, 1);
    mbar_init(&v_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);   // warp-uniform for ptxas
  const uint32_t tmem_s = tmem_base;
  pdl_wait();
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 8) {
    // whole warp runs the loop (uniform control flow), one elected lane issues: see ptx.cuh
    if (elect_one()) {
      mbar_arrive_expect_tx(&q_full, kQBytes);
      tma_load_3d(q_s, &tmQ, &q_full, h * kD, t0, b);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[s], kKVBytes);
        tma_load_3d(k_s + s * kKVBytes, &tmK, &k_full[s], h * kD, j * kKT, b);
      }
      __syncwarp();
      mbar_wait(&v_empty, (j & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full, kKVBytes);
        tma_load_3d(v_s, &tmV, &v_full, h * kD, j * kKT, b);
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    constexpr uint32_t idesc_qk = umma_idesc_bf16(kQT, kKT, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(kQT, kD, 0, 1);   // B (= V) is MN-major
    mbar_wait(&q_full, 0);
    const uint64_t q_desc = umma_desc_kmajor<128>(smem_u32(q_s));
    const uint64_t p_desc = umma_desc_kmajor<128>(smem_u32(p_s));
    const uint64_t v_desc = umma_desc_mnmajor_sw128(smem_u32(v_s), 1024);
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      mbar_wait(&k_full[s], (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t k_desc = umma_desc_kmajor<128>(smem_u32(k_s + s * kKVBytes));
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk)
          umma_bf16(tmem_s, q_desc + ((kk * 32) >> 4), k_desc + ((kk * 32) >> 4), idesc_qk, kk != 0);
        umma_commit(&s_full);
        umma_commit(&k_empty[s]);
      }
      __syncwarp();
      mbar_wait(&p_full, j & 1);
      mbar_wait(&v_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kKT / 16; ++kk)
          umma_bf16(tmem_o, p_desc + (((kk >> 2) * (kQT * 128) + (kk & 3) * 32) >> 4),
                    v_desc + ((kk * 16 * 128) >> 4), idesc_pv, (j | kk) != 0);
        umma_commit(&v_empty);
        if (j == n_tiles - 1) umma_commit(&o_done);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------- softmax / epilogue
    const int half = warp >> 2;                  // which 64 keys of the tile / 32 columns of O
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full, j & 1);
      tc_fence_after();
      const int key0 = j * kKT + half * 64;
      uint32_t r[64];
      {
        uint32_t (&r0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[0]);
        uint32_t (&r1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[32]);
        tmem_ld32(tmem_s + lane_off + half * 64, r0);
        tmem_ld32(tmem_s + lane_off + half * 64 + 32, r1);
        tmem_ld_wait();
      }
      float m_loc = -INFINITY;
      if (key0 + 64 <= p.Tk) {
#pragma unroll
        for (int i = 0; i < 64; ++i) m_loc = fmaxf(m_loc, __uint_as_float(r[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          if (key0 + i >= p.Tk) r[i] = 0xff800000u;   // -inf: key outside the sequence
          m_loc = fmaxf(m_loc, __uint_as_float(r[i]));
        }
      }
      s_xch[half][row] = m_loc;
      named_bar_sync(2, kSoftmaxThreads);
      const float m_new = fmaxf(m_run, fmaxf(m_loc, s_xch[half ^ 1][row]));
      const float alpha = ex2_approx((m_run - m_new) * p.scale_log2);   // 0 on the first tile
      const float m_scaled = m_new * p.scale_log2;
      // rescale the running output only if some row of this warp moved its max (the previous
      // P V is complete: s_full is committed after it)
      if (j > 0 && __any_sync(0xffffffffu, m_new > m_run)) {
        uint32_t o[32];
        tmem_ld32(tmem_o + lane_off + half * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        uint32_t (&o0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&o[0]);
        uint32_t (&o1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&o[16]);
        tmem_st16(tmem_o + lane_off + half * 32, o0);
        tmem_st16(tmem_o + lane_off + half * 32 + 16, o1);
        tmem_st_wait();
      }
      // p = 2^(s*c - m*c) -> bf16 -> swizzled smem chunk `half` (K-major A operand of P V)
      float l_tile = 0.f;
      uint8_t* prow = p_s + half * (kQT * 128) + row * 128;
#pragma unroll
      for (int piece = 0; piece < 8; ++piece) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p0 = ex2_approx(__uint_as_float(r[piece * 8 + 2 * i]) * p.scale_log2 - m_scaled);
          const float p1 = ex2_approx(__uint_as_float(r[piece * 8 + 2 * i + 1]) * p.scale_log2 - m_scaled);
          pk[i] = pack_bf16(p0, p1);
          const float2 pr = unpack_bf16(pk[i]);   // sum what the MMA will actually see
          l_tile += pr.x + pr.y;
        }
        *reinterpret_cast<uint4*>(prow + ((piece ^ (row & 7)) << 4)) =
            make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full);
    }
    // total row sum = sum of the two halves
    mbar_wait(&o_done, 0);
    tc_fence_after();
    s_xch[half][row] = l_run;
    named_bar_sync(2, kSoftmaxThreads);
    const float l_tot = l_run + s_xch[half ^ 1][row];
    const float inv_l = 1.f / l_tot;
    const int t = t0 + row;
    if (p.lse != nullptr && half == 0 && t < p.Tq)    // natural-log units (adp_attention_bwd)
      p.lse[(static_cast<size_t>(b) * gridDim.y + h) * p.Tq + t] =
          (m_run * p.scale_log2 + log2f(l_tot)) * 0.6931471805599453f;
    uint32_t o[32];
    tmem_ld32(tmem_o + lane_off + half * 32, o);
    tmem_ld_wait();
    if (t < p.Tq) {
      __nv_bfloat16* orow = p.o + (static_cast<size_t>(b) * p.Tq + t) * p.ldo + h * kD + half * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        uint4 ov;
        ov.x = pack_bf16(__uint_as_float(o[c]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
        ov.y = pack_bf16(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
        ov.z = pack_bf16(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
        ov.w = pack_bf16(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
        *reinterpret_cast<uint4*>(orow + c) = ov;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace adp

using namespace adp;

extern "C" int adp_attention(const void* q, const void* k, const void* v, void* o, int32_t B,
                             int32_t H, int32_t Tq, int32_t Tk, int32_t ldq, int32_t ldk,
                             int32_t ldv, int32_t ldo, float scale, float* lse,
                             adp_stream_t stream) {
  ADP_CHECK(q && k && v && o, "adp_attention: null pointer");
  ADP_CHECK(B > 0 && H > 0 && Tq > 0 && Tk > 0, "adp_attention: bad sizes");
  ADP_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && ldq >= H * kD &&
                ldk >= H * kD && ldv >= H * kD && ldo >= H * kD,
            "adp_attention: row pitches must be multiples of 8 and >= heads*64");
  ADP_CHECK(scale > 0.f, "adp_attention: scale must be positive");
  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[3] = {(uint32_t)kD, (uint32_t)kQT, 1};
  {
    const uint64_t dims[3] = {(uint64_t)H * kD, (uint64_t)Tq, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)Tq * ldq * 2};
    if (int e = make_tmap_bf16(&tmQ, q, 3, dims, str, box, 128)) return e;
  }
  {
    const uint64_t dims[3] = {(uint64_t)H * kD, (uint64_t)Tk, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)ldk * 2, (uint64_t)Tk * ldk * 2};
    if (int e = make_tmap_bf16(&tmK, k, 3, dims, str, box, 128)) return e;
  }
  {
    const uint64_t dims[3] = {(uint64_t)H * kD, (uint64_t)Tk, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)ldv * 2, (uint64_t)Tk * ldv * 2};
    if (int e = make_tmap_bf16(&tmV, v, 3, dims, str, box, 128)) return e;
  }
  static SmemAttrCache smem_cache;
  ADP_CUDA(ensure_dyn_smem(attention_kernel, (size_t)kAttnSmem, smem_cache));
  AttnParams p;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.lse = lse;
  p.Tq = Tq;
  p.Tk = Tk;
  p.ldo = ldo;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((Tq + kQT - 1) / kQT, H, B);
  ADP_CUDA(launch_k(attention_kernel, grid, dim3(320), (size_t)kAttnSmem, as_stream(stream), tmQ, tmK, tmV, p));
  ADP_LAUNCH_CHECK();
  return 0;
}
