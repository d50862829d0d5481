// Draft-block attention: the bs block queries of every request attend, without any mask, to
// [committed context K/V | this block's K/V] = cache positions [0, start + bs)
// (model/dflash.py:77-99: is_causal=False, attention_mask=None, GQA, scale D^-1/2).
//
// Split-KV ("flash-decoding") so that short query blocks still fill the machine: work item = (kv split,
// kv head, request x 16-query tile); the `group` q-heads that share the kv head are the item's warps,
// so each K/V byte is fetched from HBM once. Scores and PV run on mma.sync m16n8k16 (bf16 in, fp32
// accumulate): with 16 queries per head this is <1% of the step's bytes and FLOPs, far below what
// would amortise a TMEM round trip. A second tiny pass merges the splits.
//
// The bodies are __device__ functions over (work item, thread-in-group) so that both the stand-alone
// kernels and the persistent step kernel (step_mega.cuh) run the same code. Data produced earlier in
// the same step by other CTAs is read with ld.global.cg (L2), never through L1.
#pragma once
#include "ptx.cuh"

namespace dfl {

constexpr int kAttnKeys = 64;  // keys per smem tile
constexpr int kAttnD = 128;
constexpr int kAttnTileBytes = kAttnKeys * kAttnD * 2;  // 16 KB (K or V)
constexpr int kAttnSmem = 2 * 2 * kAttnTileBytes;       // stand-alone kernel: 2 stages x (K, V)

struct AttnArgs {
  int R, SL, bs, Hq, Hkv, S_max;
  int nsplit;
  const int* start;
  const int* blk_len;
  const int* ctx_len;            // optional: keys below start - ctx_len were cached by earlier steps
  const __nv_bfloat16* q;        // [R*SL][Hq][128]
  const __nv_bfloat16* k_cache;  // [R][Hkv][S_max][128]
  const __nv_bfloat16* v_cache;
  float* part_o;   // [nsplit][R*SL][Hq][128]  unnormalised
  float* part_ml;  // [nsplit][R*SL][Hq][2]    (running max in log2 domain, running sum)
  float scale_log2;  // D^-1/2 * log2(e)
  __nv_bfloat16* out;  // [R*SL][Hq*128]
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// keys handled by one split, a multiple of the smem tile
__host__ __device__ inline int attn_chunk(int L, int nsplit) {
  int c = (L + nsplit - 1) / nsplit;
  c = (c + kAttnKeys - 1) / kAttnKeys * kAttnKeys;
  return c < kAttnKeys ? kAttnKeys : c;
}

// One (split, kv head h, request r, query tile qt) work item, executed by `group` warps
// (tid in [0, 32*group)). NSTG = K/V smem stages (1 or 2) at smem0; bar_id = named barrier of the group.
// LOCAL: partials are indexed (row_in_tile * group + head_in_group) -- a per-CTA buffer (shared memory of the
// cluster-fused kernel) -- instead of the global [split][row][head] layout.
// PDLIN (stand-alone kernel): griddepcontrol.wait happens INSIDE, after the first K/V tile has been requested when
// that tile only holds keys cached by earlier steps (old_end = start - ctx_len): its HBM latency then overlaps the
// wait for qkv_post instead of following it.
template <int NSTG, bool LOCAL = false, bool PDLIN = false>
__device__ __forceinline__ void attn_split_body(const AttnArgs& a, int r, int qt, int h, int split, int tid,
                                                int nthreads, uint32_t smem0, int bar_id, int L_in = -1,
                                                int old_end = -1) {
  const int warp = tid >> 5, lane = tid & 31;
  const int group = a.Hq / a.Hkv;
  const int hq = h * group + warp;
  const int RS = a.R * a.SL;
  const int L = L_in >= 0 ? L_in : a.start[r] + a.blk_len[r];
  const int chunk = attn_chunk(L, a.nsplit);
  const int k0 = split * chunk;
  const int k1 = min(L, k0 + chunk);
  const int g = lane >> 2, tq = lane & 3;
  const int row_lo = r * a.SL + qt * 16 + g;  // this thread's two query rows: row_lo, row_lo + 8

  auto part_index = [&](int rl) -> long long {
    if (LOCAL) return static_cast<long long>(rl - (r * a.SL + qt * 16)) * group + warp;
    return (static_cast<long long>(split) * RS + rl) * a.Hq + hq;
  };
  if (k0 >= L) {
    if (tq == 0) {
      a.part_ml[part_index(row_lo) * 2 + 0] = -INFINITY;
      a.part_ml[part_index(row_lo) * 2 + 1] = 0.f;
      a.part_ml[part_index(row_lo + 8) * 2 + 0] = -INFINITY;
      a.part_ml[part_index(row_lo + 8) * 2 + 1] = 0.f;
    }
    return;
  }

  const __nv_bfloat16* kbase = a.k_cache + (static_cast<long long>(r) * a.Hkv + h) * a.S_max * kAttnD;
  const __nv_bfloat16* vbase = a.v_cache + (static_cast<long long>(r) * a.Hkv + h) * a.S_max * kAttnD;

  auto load_tile = [&](int t, int stage) {
    const int key0 = k0 + t * kAttnKeys;
    const uint32_t sk = smem0 + stage * 2 * kAttnTileBytes;
    const uint32_t sv = sk + kAttnTileBytes;
    for (int idx = tid; idx < kAttnKeys * 16; idx += nthreads) {
      const int key = idx >> 4, ch = idx & 15;
      const bool valid = key0 + key < k1;
      const long long goff = static_cast<long long>(valid ? key0 + key : 0) * kAttnD + ch * 8;
      const uint32_t so = key * 256 + ((ch ^ (key & 7)) << 4);
      cp_async16(sk + so, kbase + goff, valid);
      cp_async16(sv + so, vbase + goff, valid);
    }
  };

  bool pre0 = false;
  if (PDLIN) {
    pre0 = NSTG == 2 && old_end >= 0 && k0 + kAttnKeys <= old_end;
    if (pre0) {
      load_tile(0, 0);
      cp_async_commit();
    }
    DFL_WAIT_THEN_TRIGGER();
  }

  // Q fragments (A operand, 8 k-steps of 16 over D=128), straight from global (L2)
  uint32_t qf[8][4];
  {
    const uint32_t* q0 = reinterpret_cast<const uint32_t*>(a.q + (static_cast<long long>(row_lo) * a.Hq + hq) * kAttnD);
    const uint32_t* q1 =
        reinterpret_cast<const uint32_t*>(a.q + (static_cast<long long>(row_lo + 8) * a.Hq + hq) * kAttnD);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const int c = (ks * 16 + tq * 2) >> 1;  // in 32-bit words
      qf[ks][0] = __ldcg(q0 + c);
      qf[ks][1] = __ldcg(q1 + c);
      qf[ks][2] = __ldcg(q0 + c + 4);
      qf[ks][3] = __ldcg(q1 + c + 4);
    }
  }

  float o[16][4];
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  const int ntiles = (k1 - k0 + kAttnKeys - 1) / kAttnKeys;
  if (NSTG == 2 && !pre0) {
    load_tile(0, 0);
    cp_async_commit();
  }
  for (int t = 0; t < ntiles; ++t) {
    int stage = 0;
    if (NSTG == 2) {
      stage = t & 1;
      if (t + 1 < ntiles) {
        load_tile(t + 1, stage ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
    } else {
      load_tile(t, 0);
      cp_async_commit();
      cp_async_wait<0>();
    }
    group_sync(bar_id, nthreads);
    const uint32_t sk = smem0 + stage * 2 * kAttnTileBytes;
    const uint32_t sv = sk + kAttnTileBytes;

    // S = Q K^T : 8 n-tiles of 8 keys
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of n-tiles
        const int mtx = lane >> 3;
        const int key = np * 16 + (mtx >> 1) * 8 + (lane & 7);
        const int ch = ks * 2 + (mtx & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sk + key * 256 + ((ch ^ (key & 7)) << 4), b0, b1, b2, b3);
        mma_16816(s[np * 2], qf[ks], b0, b1);
        mma_16816(s[np * 2 + 1], qf[ks], b2, b3);
      }
    }
    // mask keys beyond the split end, online softmax (log2 domain)
    const int key_base = k0 + t * kAttnKeys;
    float tmax_lo = -INFINITY, tmax_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = key_base + nt * 8 + tq * 2 + e < k1;
        s[nt][e] = ok ? s[nt][e] * a.scale_log2 : -INFINITY;
        s[nt][2 + e] = ok ? s[nt][2 + e] * a.scale_log2 : -INFINITY;
        tmax_lo = fmaxf(tmax_lo, s[nt][e]);
        tmax_hi = fmaxf(tmax_hi, s[nt][2 + e]);
      }
    }
    tmax_lo = fmaxf(tmax_lo, __shfl_xor_sync(0xffffffffu, tmax_lo, 1));
    tmax_lo = fmaxf(tmax_lo, __shfl_xor_sync(0xffffffffu, tmax_lo, 2));
    tmax_hi = fmaxf(tmax_hi, __shfl_xor_sync(0xffffffffu, tmax_hi, 1));
    tmax_hi = fmaxf(tmax_hi, __shfl_xor_sync(0xffffffffu, tmax_hi, 2));
    const float mn_lo = fmaxf(m_lo, tmax_lo), mn_hi = fmaxf(m_hi, tmax_hi);
    const float al_lo = exp2f(m_lo - mn_lo), al_hi = exp2f(m_hi - mn_hi);  // exp2(-inf) = 0 on first tile
    m_lo = mn_lo;
    m_hi = mn_hi;
    l_lo *= al_lo;
    l_hi *= al_hi;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      o[nt][0] *= al_lo; o[nt][1] *= al_lo;
      o[nt][2] *= al_hi; o[nt][3] *= al_hi;
    }
    uint32_t pf[4][4];  // P as A operand: 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - mn_lo), p1 = exp2f(s[nt][1] - mn_lo);
      const float p2 = exp2f(s[nt][2] - mn_hi), p3 = exp2f(s[nt][3] - mn_hi);
      l_lo += p0 + p1;
      l_hi += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 8; ++dp) {  // pairs of d n-tiles
        const int mtx = lane >> 3;
        const int key = kk * 16 + (mtx & 1) * 8 + (lane & 7);
        const int ch = dp * 2 + (mtx >> 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sv + key * 256 + ((ch ^ (key & 7)) << 4), b0, b1, b2, b3);
        mma_16816(o[dp * 2], pf[kk], b0, b1);
        mma_16816(o[dp * 2 + 1], pf[kk], b2, b3);
      }
    }
    group_sync(bar_id, nthreads);
  }

  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  if (!LOCAL && a.nsplit == 1) {
    // one split per (request, kv head): nothing to merge -- normalise and write the bf16 attention output here (the
    // merge kernel is not launched). Same arithmetic as attn_combine_item with a single live split (weight exp2(0) = 1).
    const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
    __nv_bfloat16* d_lo = a.out + (static_cast<long long>(row_lo) * a.Hq + hq) * kAttnD;
    __nv_bfloat16* d_hi = a.out + (static_cast<long long>(row_lo + 8) * a.Hq + hq) * kAttnD;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const int d = nt * 8 + tq * 2;
      *reinterpret_cast<uint32_t*>(d_lo + d) = pack_bf16(o[nt][0] * inv_lo, o[nt][1] * inv_lo);
      *reinterpret_cast<uint32_t*>(d_hi + d) = pack_bf16(o[nt][2] * inv_hi, o[nt][3] * inv_hi);
    }
    return;
  }
  const long long p_lo = part_index(row_lo), p_hi = part_index(row_lo + 8);
  if (tq == 0) {
    a.part_ml[p_lo * 2 + 0] = m_lo; a.part_ml[p_lo * 2 + 1] = l_lo;
    a.part_ml[p_hi * 2 + 0] = m_hi; a.part_ml[p_hi * 2 + 1] = l_hi;
  }
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
    const int d = nt * 8 + tq * 2;
    *reinterpret_cast<float2*>(a.part_o + p_lo * kAttnD + d) = make_float2(o[nt][0], o[nt][1]);
    *reinterpret_cast<float2*>(a.part_o + p_hi * kAttnD + d) = make_float2(o[nt][2], o[nt][3]);
  }
}

__global__ void attn_split_kernel(const AttnArgs a) {
  extern __shared__ __align__(128) uint8_t attn_smem[];
  const int tiles_per_req = a.SL / 16;
  const int r = blockIdx.z / tiles_per_req;
  // the key range comes from request state that only the previous step's accept kernel writes: read it while the
  // kernel in front of this one (qkv_post) is still running
  const int st = a.start[r];
  const int L = st + a.blk_len[r];
  const int old_end = a.ctx_len != nullptr ? st - a.ctx_len[r] : -1;
  attn_split_body<2, false, true>(a, r, blockIdx.z % tiles_per_req, blockIdx.y, blockIdx.x, threadIdx.x, blockDim.x,
                                  smem_u32(attn_smem), 0, L, old_end);
  DFL_TRACE(2);
}

// Merge the splits of one (row, q head) item: one warp.
__device__ __forceinline__ void attn_combine_item(const AttnArgs& a, int item, int lane) {
  const int RS = a.R * a.SL;
  // all (max, sum) pairs first (one L2 round trip), then the partial outputs of the live splits
  // (max, sum) pairs AND the partial outputs of every split are requested together: one L2 round trip instead of two
  // (a dead split's output rows were never written; they are loaded but only used under the `!= -inf` test below)
  float2 ml[16];
  float4 ov[16];
#pragma unroll
  for (int s = 0; s < 16; ++s)
    if (s < a.nsplit) {
      ml[s] = __ldcg(reinterpret_cast<const float2*>(a.part_ml + (static_cast<long long>(s) * RS * a.Hq + item) * 2));
      ov[s] = __ldcg(reinterpret_cast<const float4*>(a.part_o + (static_cast<long long>(s) * RS * a.Hq + item) * kAttnD + lane * 4));
    }
  float M = -INFINITY;
#pragma unroll
  for (int s = 0; s < 16; ++s)
    if (s < a.nsplit) M = fmaxf(M, ml[s].x);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float den = 0.f;
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    if (s < a.nsplit && ml[s].x != -INFINITY) {
      const float w = exp2f(ml[s].x - M);
      den += w * ml[s].y;
      acc[0] += w * ov[s].x; acc[1] += w * ov[s].y; acc[2] += w * ov[s].z; acc[3] += w * ov[s].w;
    }
  }
  const float inv = 1.0f / den;
  __nv_bfloat16* dst = a.out + static_cast<long long>(item) * kAttnD + lane * 4;  // item = row*Hq + head
  *reinterpret_cast<uint2*>(dst) =
      make_uint2(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv));
}

__global__ void __launch_bounds__(32 * kItemWarps) attn_combine_kernel(const AttnArgs a) {
  DFL_WAIT_THEN_TRIGGER();
  const int item = blockIdx.x * kItemWarps + (threadIdx.x >> 5);
  if (item >= a.R * a.SL * a.Hq) return;
  attn_combine_item(a, item, threadIdx.x & 31);
  DFL_TRACE(2);
}

}  // namespace dfl
