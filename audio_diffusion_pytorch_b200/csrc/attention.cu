// adp_attention: softmax(q k^T * scale) v, head dim 64, on tcgen05 (a_unet AttentionBase).
//
// One CTA = one (batch, head, 128-query tile).  S = Q K^T (128 x 128 keys) and O (128 x 64)
// accumulate in TMEM; the 128 softmax threads own one query row each (TMEM lane == row), so
// the online-softmax max / sum / rescale are thread-local, no shuffles.
//   warp 4: TMA producer   Q once, then a 2-slot ring of {K tile, V tile}
//   warp 5: MMA issuer     S = Q K_j^T ; (wait P) ; O += P_j V_j
//   warps 0-3: softmax     S(TMEM) -> max/exp2/sum -> P(bf16, swizzled smem) ; rescale O(TMEM)
// K is the K-major B operand of the first GEMM (keys x d); V is used in place as the
// MN-major B operand of the second (d contiguous per key), so no transpose is materialised.
// Two CTAs fit per SM (112 KB smem, 256 TMEM columns each): one CTA's softmax overlaps the
// other's MMAs.
#include "common.cuh"
#include "ptx.cuh"

namespace adp {

constexpr int kQT = 128;       // query rows per CTA
constexpr int kKT = 128;       // keys per tile
constexpr int kD = 64;         // head dim
constexpr int kQBytes = kQT * kD * 2;          // 16 KB
constexpr int kKVBytes = kKT * kD * 2;         // 16 KB
constexpr int kPBytes = kQT * kKT * 2;         // 32 KB (two 64-key swizzle chunks)
constexpr int kAttnSmem = kQBytes + 2 * 2 * kKVBytes + kPBytes + 1024;
constexpr uint32_t kTmemCols = 256;            // S: [0,128)  O: [128,192)

struct AttnParams {
  __nv_bfloat16* o;
  int Tq, Tk, ldo;
  float scale_log2;   // scale * log2(e)
};

__global__ void __launch_bounds__(192)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, s_full, p_full, o_done;
  __shared__ uint64_t kv_full[2], kv_empty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* q_s = base;
  uint8_t* k_s = q_s + kQBytes;            // [2][16 KB]
  uint8_t* v_s = k_s + 2 * kKVBytes;       // [2][16 KB]
  uint8_t* p_s = v_s + 2 * kKVBytes;       // 32 KB

  const int t0 = blockIdx.x * kQT;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (p.Tk + kKT - 1) / kKT;

  if (warp == 4) {
    tmem_alloc(&tmem_slot, kTmemCols);
    tmem_relinquish();
  } else if (warp == 5 && lane == 0) {
    mbar_init(&q_full, 1);
    mbar_init(&s_full, 1);
    mbar_init(&p_full, 128);
    mbar_init(&o_done, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&q_full, kQBytes);
      tma_load_3d(q_s, &tmQ, &q_full, h * kD, t0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * kKVBytes);
        tma_load_3d(k_s + s * kKVBytes, &tmK, &kv_full[s], h * kD, j * kKT, b);
        tma_load_3d(v_s + s * kKVBytes, &tmV, &kv_full[s], h * kD, j * kKT, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kQT, kKT, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kQT, kD, 0, 1);   // B (= V) is MN-major
      mbar_wait(&q_full, 0);
      const uint32_t q_addr = smem_u32(q_s);
      const uint32_t p_addr = smem_u32(p_s);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(k_s + s * kKVBytes);
        const uint32_t v_addr = smem_u32(v_s + s * kKVBytes);
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk)
          umma_bf16(tmem_s, umma_desc_kmajor<128>(q_addr + kk * 32),
                    umma_desc_kmajor<128>(k_addr + kk * 32), idesc_qk, kk != 0);
        umma_commit(&s_full);
        mbar_wait(&p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kKT / 16; ++kk)
          umma_bf16(tmem_o,
                    umma_desc_kmajor<128>(p_addr + (kk >> 2) * (kQT * 128) + (kk & 3) * 32),
                    umma_desc_mnmajor_sw128(v_addr + kk * 16 * 128, 1024), idesc_pv,
                    (j | kk) != 0);
        umma_commit(&kv_empty[s]);
      }
      umma_commit(&o_done);
    }
  } else {
    // ------------------------------------------------------------- softmax / epilogue
    const int row = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full, j & 1);
      tc_fence_after();
      const int key0 = j * kKT;
      // pass 1: row max
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kKT; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_off + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sv = (key0 + c + i < p.Tk) ? __uint_as_float(r[i]) : -INFINITY;
          m_tile = fmaxf(m_tile, sv);
        }
      }
      const float m_new = fmaxf(m_run, m_tile);
      const float alpha = exp2f((m_run - m_new) * p.scale_log2);   // 0 on the first tile
      const float m_scaled = m_new * p.scale_log2;
      if (j > 0) {   // rescale the running output (previous P V already complete: see s_full)
#pragma unroll 1
        for (int c = 0; c < kD; c += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_o + lane_off + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st16(tmem_o + lane_off + c, r);
        }
        tmem_st_wait();
      }
      // pass 2: p = exp2(s*c - m*c) -> bf16 -> swizzled smem (K-major A operand of P V)
      float l_tile = 0.f;
#pragma unroll 1
      for (int c = 0; c < kKT; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_off + c, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = 0.f, p1 = 0.f;
          if (key0 + c + i < p.Tk) p0 = exp2f(__uint_as_float(r[i]) * p.scale_log2 - m_scaled);
          if (key0 + c + i + 1 < p.Tk)
            p1 = exp2f(__uint_as_float(r[i + 1]) * p.scale_log2 - m_scaled);
          pk[i >> 1] = pack_bf16(p0, p1);
          const float2 pr = unpack_bf16(pk[i >> 1]);   // sum what the MMA will actually see
          l_tile += pr.x + pr.y;
        }
        // 32 keys = four 16-byte pieces of this row inside 64-key chunk (c / 64)
        uint8_t* chunk = p_s + (c >> 6) * (kQT * 128) + row * 128;
        const int piece0 = (c & 63) >> 3;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int piece = (piece0 + q4) ^ (row & 7);
          *reinterpret_cast<uint4*>(chunk + piece * 16) =
              make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        }
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full);
    }
    mbar_wait(&o_done, 0);
    tc_fence_after();
    const int t = t0 + row;
    const float inv_l = 1.f / l_run;
    __nv_bfloat16* orow = p.o + (static_cast<size_t>(b) * p.Tq + (t < p.Tq ? t : 0)) * p.ldo + h * kD;
#pragma unroll 1
    for (int c = 0; c < kD; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_o + lane_off + c, r);
      tmem_ld_wait();
      if (t < p.Tq) {
        uint4 o0, o1;
        o0.x = pack_bf16(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l);
        o0.y = pack_bf16(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l);
        o0.z = pack_bf16(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l);
        o0.w = pack_bf16(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l);
        o1.x = pack_bf16(__uint_as_float(r[8]) * inv_l, __uint_as_float(r[9]) * inv_l);
        o1.y = pack_bf16(__uint_as_float(r[10]) * inv_l, __uint_as_float(r[11]) * inv_l);
        o1.z = pack_bf16(__uint_as_float(r[12]) * inv_l, __uint_as_float(r[13]) * inv_l);
        o1.w = pack_bf16(__uint_as_float(r[14]) * inv_l, __uint_as_float(r[15]) * inv_l);
        *reinterpret_cast<uint4*>(orow + c) = o0;
        *reinterpret_cast<uint4*>(orow + c + 8) = o1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace adp

using namespace adp;

extern "C" int adp_attention(const void* q, const void* k, const void* v, void* o, int32_t B,
                             int32_t H, int32_t Tq, int32_t Tk, int32_t ldq, int32_t ldk,
                             int32_t ldv, int32_t ldo, float scale, adp_stream_t stream) {
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
  static bool attr_set = false;
  if (!attr_set) {
    ADP_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kAttnSmem));
    attr_set = true;
  }
  AttnParams p;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.Tq = Tq;
  p.Tk = Tk;
  p.ldo = ldo;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((Tq + kQT - 1) / kQT, H, B);
  attention_kernel<<<grid, 192, kAttnSmem, as_stream(stream)>>>(tmQ, tmK, tmV, p);
  ADP_LAUNCH_CHECK();
  return 0;
}
