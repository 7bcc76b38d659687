// Software-pipelined epilogue of the fused GRU kernels (bf16 training instantiations of EPI_ZR, EPI_H, EPI_DH, EPI_DRH).
//
// A tile is drained in 32-column chunks.  Each chunk's operands live in one "set" of shared memory:
//   fp32 tensors as [32 rows x 128 B] SWIZZLE_128B boxes, bf16 tensors as [32 rows x 64 B] SWIZZLE_64B boxes.
// Two sets ping-pong: the TMA loads of the NEXT chunk (of this tile, or of the first chunk of this warp's next tile)
// are issued before the current chunk is processed, so the load latency is off the critical path; outputs overwrite
// their inputs in place where shapes match, and leave through TMA stores.
#pragma once

namespace srg {

// 16-byte chunk c (0..3) of row r inside a [32 x 64 B] SWIZZLE_64B box: bits [4,5] ^= bits [7,8]
__device__ __forceinline__ uint32_t u64_addr(uint32_t base, int r, int c) {
  return base + r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
}
__device__ __forceinline__ void u64_ld_bf16x8(uint32_t base, int lane, int c, float (&v)[8]) {
  uint32_t u[4];
  lds128(u64_addr(base, lane, c), u[0], u[1], u[2], u[3]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = bf16_lo_f(u[i]);
    v[2 * i + 1] = bf16_hi_f(u[i]);
  }
}
__device__ __forceinline__ void u64_st_bf16x8(uint32_t base, int lane, int c, const float (&v)[8]) {
  sts128(u64_addr(base, lane, c), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
         pack_bf16(v[6], v[7]));
}

template <int EPI>
struct PipeTraits {
  // bytes of one input/output set per warp
  static constexpr int kSetBytes = (EPI == EPI_ZR) ? 8192 : (EPI == EPI_H) ? 8192 : (EPI == EPI_DH) ? 12288 : 4096;
};

constexpr int kF32Box = 4096;   // [32 x 32] fp32
constexpr int kB16Box = 2048;   // [32 x 32] bf16

// Everything one epilogue warp needs to know about its place in the kernel.
struct PipeCtx {
  const GemmMaps* maps;
  const GemmArgs* args;
  uint8_t* smem;       // this warp's epilogue region (2 sets)
  uint64_t* in_bar;    // [2]
  int lane;
};

// ---- per-EPI: issue the loads of one chunk into set `sp` (lane 0 only) -------------------------------------------
template <int EPI>
__device__ __forceinline__ void pipe_issue(const PipeCtx& c, uint8_t* sp, uint64_t* bar, int n0, int cc, int row0) {
  const GemmMaps& m = *c.maps;
  const GemmArgs& a = *c.args;
  const int gcol = n0 + cc * 32;
  if constexpr (EPI == EPI_ZR) {
    // r-tiles only: io1 = h fp32
    ptx::mbar_arrive_expect_tx(bar, kF32Box);
    ptx::tma_load_2d(&m.io[1], bar, sp, gcol - a.n_split, row0);
  } else if constexpr (EPI == EPI_H) {
    ptx::mbar_arrive_expect_tx(bar, kF32Box + kB16Box);
    ptx::tma_load_2d(&m.io[0], bar, sp, gcol, row0);                 // h fp32
    ptx::tma_load_2d(&m.io[1], bar, sp + 4096, gcol, row0);          // z bf16
  } else if constexpr (EPI == EPI_DH) {
    const bool next = (a.flags & FLAG_NEXT) != 0, has_add = (a.flags & FLAG_ADD) != 0;
    ptx::mbar_arrive_expect_tx(bar, kF32Box + (next ? 3 : 0) * kB16Box + (has_add ? kB16Box : 0));
    ptx::tma_load_2d(&m.io[0], bar, sp, gcol, row0);                 // dh_acc fp32
    if (next) {
      ptx::tma_load_2d(&m.io[1], bar, sp + 4096, gcol, row0);        // z
      ptx::tma_load_2d(&m.io[2], bar, sp + 6144, gcol, row0);        // hc
      ptx::tma_load_2d(&m.io[3], bar, sp + 8192, gcol, row0);        // h
    }
    if (has_add) ptx::tma_load_2d(&m.io[6], bar, sp + 10240, gcol, row0);
  } else {  // EPI_DRH
    ptx::mbar_arrive_expect_tx(bar, 2 * kB16Box);
    ptx::tma_load_2d(&m.io[0], bar, sp, gcol, row0);                 // h
    ptx::tma_load_2d(&m.io[1], bar, sp + 2048, gcol, row0);          // r
  }
}

// ---- per-EPI: compute one chunk (all lanes) and write the results into the set -----------------------------------
template <int EPI>
__device__ __forceinline__ void pipe_compute(const PipeCtx& c, uint32_t s, const float (&acc)[32], int n0, int cc,
                                             bool r_tile) {
  const GemmArgs& a = *c.args;
  const int lane = c.lane;
  const int gcol = n0 + cc * 32;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if constexpr (EPI == EPI_ZR) {
      float b[8];
      load_bias8(a.bias, gcol + g * 8, 1.0f, b);
      if (!r_tile) {   // z = sigmoid(acc + b) -> bf16 at offset 0
        float z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = act_sigmoid<false>(acc[g * 8 + i] + b[i]);
        u64_st_bf16x8(s, lane, g, z);
      } else {         // r = sigmoid(acc + b); rh = r * h
        float h[8], r[8], rh[8];
        slot_ld_f32x8(s, lane, g, h);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          r[i] = act_sigmoid<false>(acc[g * 8 + i] + b[i]);
          rh[i] = r[i] * h[i];
        }
        u64_st_bf16x8(s + 4096, lane, g, rh);
        if (a.flags & FLAG_STASH) u64_st_bf16x8(s + 6144, lane, g, r);
      }
    } else if constexpr (EPI == EPI_H) {
      float b[8], h[8], z[8], hc[8], hn[8];
      load_bias8(a.bias, gcol + g * 8, 1.0f, b);
      slot_ld_f32x8(s, lane, g, h);
      u64_ld_bf16x8(s + 4096, lane, g, z);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        hc[i] = act_tanh<false>(acc[g * 8 + i] + b[i]);
        hn[i] = fmaf(z[i], hc[i] - h[i], h[i]);
      }
      slot_st_f32x8(s, lane, g, hn);                                   // h' fp32, in place
      u64_st_bf16x8(s + 6144, lane, g, hn);                            // h' bf16 operand copy
      if (a.flags & FLAG_STASH) u64_st_bf16x8(s + 4096, lane, g, hc);  // hc over z
    } else if constexpr (EPI == EPI_DH) {
      const bool next = (a.flags & FLAG_NEXT) != 0, has_add = (a.flags & FLAG_ADD) != 0;
      float dh[8];
      slot_ld_f32x8(s, lane, g, dh);
#pragma unroll
      for (int i = 0; i < 8; ++i) dh[i] += acc[g * 8 + i];
      if (has_add) {
        float ad[8];
        u64_ld_bf16x8(s + 10240, lane, g, ad);
#pragma unroll
        for (int i = 0; i < 8; ++i) dh[i] += ad[i];
      }
      if (next) {
        float z[8], hc[8], h[8], dz[8], dc[8];
        u64_ld_bf16x8(s + 4096, lane, g, z);
        u64_ld_bf16x8(s + 6144, lane, g, hc);
        u64_ld_bf16x8(s + 8192, lane, g, h);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dz[i] = dh[i] * (hc[i] - h[i]) * z[i] * (1.0f - z[i]);
          dc[i] = dh[i] * z[i] * (1.0f - hc[i] * hc[i]);
          dh[i] = dh[i] * (1.0f - z[i]);
        }
        u64_st_bf16x8(s + 4096, lane, g, dz);
        u64_st_bf16x8(s + 6144, lane, g, dc);
      }
      slot_st_f32x8(s, lane, g, dh);
    } else {  // EPI_DRH
      float h[8], r[8], e[8], dp[8];
      u64_ld_bf16x8(s, lane, g, h);
      u64_ld_bf16x8(s + 2048, lane, g, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = acc[g * 8 + i];
        dp[i] = d * h[i] * r[i] * (1.0f - r[i]);
        e[i] = d * r[i];
      }
      u64_st_bf16x8(s, lane, g, dp);
      u64_st_bf16x8(s + 2048, lane, g, e);
    }
  }
}

// ---- per-EPI: TMA stores of one chunk (lane 0 only) ---------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void pipe_store(const PipeCtx& c, uint8_t* sp, int n0, int cc, int row0, bool r_tile) {
  const GemmMaps& m = *c.maps;
  const GemmArgs& a = *c.args;
  const int gcol = n0 + cc * 32;
  if constexpr (EPI == EPI_ZR) {
    if (!r_tile) {
      ptx::tma_store_2d(&m.io[0], sp, gcol, row0);
    } else {
      ptx::tma_store_2d(&m.io[2], sp + 4096, gcol - a.n_split, row0);
      if (a.flags & FLAG_STASH) ptx::tma_store_2d(&m.io[4], sp + 6144, gcol - a.n_split, row0);
    }
  } else if constexpr (EPI == EPI_H) {
    ptx::tma_store_2d(&m.io[0], sp, gcol, row0);
    ptx::tma_store_2d(&m.io[2], sp + 6144, gcol, row0);
    if (a.flags & FLAG_STASH) ptx::tma_store_2d(&m.io[4], sp + 4096, gcol, row0);
  } else if constexpr (EPI == EPI_DH) {
    ptx::tma_store_2d(&m.io[0], sp, gcol, row0);
    if (a.flags & FLAG_NEXT) {
      ptx::tma_store_2d(&m.io[4], sp + 4096, gcol, row0);
      ptx::tma_store_2d(&m.io[5], sp + 6144, gcol, row0);
    }
  } else {
    ptx::tma_store_2d(&m.io[2], sp, gcol, row0);
    ptx::tma_store_2d(&m.io[3], sp + 2048, gcol, row0);
  }
}

}  // namespace srg
