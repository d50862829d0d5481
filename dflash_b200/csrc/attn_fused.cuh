// QKV post-processing + split-KV attention + split merge in ONE kernel, built on thread-block clusters.
//
// The three stand-alone kernels (qkv_post -> attn_split -> attn_combine) cost ~15 us per layer, nearly all of it
// kernel-boundary latency. Their dependencies are local to one (request, kv head): the K/V rows and
// the `group` q heads of that kv head. So the 8 KV-split CTAs of one (request, kv head, 16-query tile) form a
// cluster and synchronise among themselves instead of through the grid:
//   A. the cluster's 8*group warps share the head's post-processing items (q/k: per-head RMSNorm + RoPE, v: round;
//      K/V rows go to the cache, q to the query buffer)                      -> barrier.cluster
//   B. each CTA runs flash-decoding over its key range; its partial (max, sum, O) stays in ITS shared memory
//                                                                            -> barrier.cluster
//   C. each CTA merges 1/8 of the (query, head) pairs, reading the 8 partials through distributed shared memory
//      (ld.shared::cluster), and writes the bf16 attention output            -> barrier.cluster (smem lifetime)
// Replaces model/dflash.py:70-99 for the block rows plus the new context rows of the cycle.
#pragma once
#include "attention.cuh"
#include "fused_ops.cuh"

namespace dfl {

constexpr int kFusedSplits = 8;  // cluster size (portable maximum)

struct AttnFusedArgs {
  QkvPostArgs post;
  AttnArgs attn;  // nsplit must be kFusedSplits; part_o / part_ml unused
  int fuse_post;  // 0: qkv_post ran as its own kernel before this one (steps B + C only)
};

__host__ __device__ inline int attn_fused_smem(int group) {
  return 2 * 2 * kAttnTileBytes + 16 * group * kAttnD * 4 + 16 * group * 2 * 4;
}

constexpr int kFusedPostWarps = 8;  // extra warps that only help with step A (latency-bound, one item per warp)

// grid (kFusedSplits, Hkv, R * SL/16), cluster (kFusedSplits, 1, 1), block 32 * (group + kFusedPostWarps)
__global__ void __launch_bounds__(384, 1) attn_fused_kernel(const AttnFusedArgs fa) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t fused_smem[];
  const QkvPostArgs& pa = fa.post;
  const AttnArgs& a = fa.attn;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int group = a.Hq / a.Hkv;
  const int nthreads = 32 * group;
  const int rank = static_cast<int>(cluster_cta_rank());
  const int h = blockIdx.y;
  const int tiles_per_req = a.SL / 16;
  const int r = blockIdx.z / tiles_per_req, qt = blockIdx.z % tiles_per_req;
  const int RS = a.R * a.SL;
  float* s_o = reinterpret_cast<float*>(fused_smem + 2 * 2 * kAttnTileBytes);  // [16*group][128]
  float* s_ml = s_o + 16 * group * kAttnD;                                    // [16*group][2]

  // ---- A: post-processing items of this (request, kv head, query tile)
  if (fa.fuse_post) {
    const int n_q = group * 16, n_kv = 2 * a.SL;  // q rows of the tile; ctx + block rows of the request
    const int total = n_q + 2 * n_kv;
    const int wpc = static_cast<int>(blockDim.x >> 5);  // warps per CTA taking part in step A
    for (int it = rank * wpc + warp; it < total; it += kFusedSplits * wpc) {
      int row, hh;
      if (it < n_q) {
        hh = h * group + it / 16;                        // q column block = q head
        row = RS + r * a.SL + qt * 16 + it % 16;         // block row
      } else {
        const int j = (it - n_q) % n_kv;
        hh = (it - n_q) < n_kv ? a.Hq + h : a.Hq + a.Hkv + h;  // k then v column block
        row = j < a.SL ? r * a.SL + j : RS + r * a.SL + (j - a.SL);
      }
      qkv_post_rowhead(pa, row, hh, lane);
    }
    __threadfence();
    cluster_sync_all();
  }

  // ---- B: flash-decoding over this CTA's key range; partial stays in shared memory
  {
    AttnArgs la = a;
    la.part_o = s_o;
    la.part_ml = s_ml;
    if (warp < group) attn_split_body<2, true>(la, r, qt, h, rank, tid, nthreads, smem_u32(fused_smem), 1);
  }
  cluster_sync_all();

  // ---- C: merge 1/8 of the (query row, head) pairs through distributed shared memory
  {
    const int pairs = 16 * group;                       // local index li = row_in_tile * group + head_in_group
    const int per_cta = (pairs + kFusedSplits - 1) / kFusedSplits;
    const uint32_t o_base = smem_u32(s_o), ml_base = smem_u32(s_ml);
    for (int k = warp; k < per_cta && warp < group; k += group) {
      const int li = rank * per_cta + k;
      if (li >= pairs) break;
      float2 ml[kFusedSplits];
      float M = -INFINITY;
#pragma unroll
      for (int s = 0; s < kFusedSplits; ++s) {
        ml[s] = dsmem_ld_f2(dsmem_map(ml_base + li * 8, s));
        M = fmaxf(M, ml[s].x);
      }
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      float den = 0.f;
#pragma unroll
      for (int s = 0; s < kFusedSplits; ++s) {
        if (ml[s].x == -INFINITY) continue;
        const float w = exp2f(ml[s].x - M);
        den += w * ml[s].y;
        const float4 ov = dsmem_ld_f4(dsmem_map(o_base + (li * kAttnD + lane * 4) * 4, s));
        acc[0] += w * ov.x; acc[1] += w * ov.y; acc[2] += w * ov.z; acc[3] += w * ov.w;
      }
      const float inv = 1.0f / den;
      const int row = r * a.SL + qt * 16 + li / group;
      const int hq = h * group + li % group;
      __nv_bfloat16* dst = a.out + (static_cast<long long>(row) * a.Hq + hq) * kAttnD + lane * 4;
      *reinterpret_cast<uint2*>(dst) =
          make_uint2(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv));
    }
  }
  cluster_sync_all();  // peers may still be reading this CTA's shared memory
}

}  // namespace dfl
