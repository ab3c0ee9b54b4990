// tcgen05 / TMA implicit-GEMM kernels for sm_100a (see igemm.cuh for the operand model).
//
// Warp roles (one CTA per SM):
//   warp 0      TMA producer   (lane 0 issues; whole warp walks the pipeline)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma / tcgen05.commit)
//   warps 2..   epilogue       (TMEM -> registers -> swizzled smem -> TMA store / red.global); the pixel GEMMs have one or
//               two groups of four epilogue warps
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue,
// double-buffered accumulators), persistent static tile schedule.
#include "igemm.cuh"

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace cs {

static constexpr int kThreads = 192;
static constexpr int kAtomBytes = 1024;            // 8 rows x 128 B: one 128B-swizzle atom
static constexpr int kASlotBytes = 18 * kAtomBytes; // (16 + 2 halo) image rows of 8 pixels x 64 ch
static constexpr int kStageBytes = 128 * 128;      // epilogue staging: 128 pixels x 64 ch bf16

// Opt a kernel in to > 48 KB of dynamic shared memory once per device (bit mask of devices already done).
template <typename K>
static cudaError_t ensure_dynamic_smem(K kern, int bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}

CS_DEVINL void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// =============================================================================================
//              BatchNorm + ReLU of an activation patch, in place in shared memory
// =============================================================================================
// The consumer of a convolution's RAW output applies the producer's BatchNorm + ReLU itself: the post-activation tensor
// is never written to (or read from) HBM.  `slot` holds `nrows` pixel rows of 64 channels (128 bytes each, 128B-swizzled
// by TMA: 16-byte chunk j of row r sits at ((j ^ (r & 7)) << 4)), row r = pixel (h_first + r / PW, w_first + r % PW).
// NT threads (a multiple of 64) sweep NT / 8 rows at a time; a thread keeps its channel chunk, so its coefficients live
// in registers and its swizzled chunk position is constant.  The arithmetic is bn_relu_kernel's: fmaxf(fmaf(y, sc, sh), 0)
// rounded to bf16 — the backward pass recomputes the same mask from y.  Pixels outside the image are set to zero (TMA
// zero-fills them in y space, but the convolution's padding is zero in activation space).
// Eight channels: two packed fp32 FMAs per 32-bit word pair (fma.rn.f32x2, sm_100) and one cvt.rn.relu.bf16x2.f32 per
// word — the same values as fmaxf(fmaf(y, sc, sh), 0) rounded to nearest-even.
struct BnCoef8 { uint64_t sc[4], sh[4]; };                  // (sc[2j], sc[2j+1]) and (sh[2j], sh[2j+1]) as f32x2
// N vectors (16 bytes = 8 channels each) at p + i * STRIDE: all loads, then all unpacks, all FMAs, all packs, all stores
// — written stage by stage so that the N x 4 independent word chains are interleaved: the warps that run this have a
// scheduler to themselves, so every dependent instruction pair costs its full pipeline latency (ncu on the first version,
// one chain per vector with a branch in between: 6 cycles per instruction, 190 cycles per vector, and the MMA warp
// waiting for the transform a third of the time).  mask[i] = 0 zeroes vector i (pixel outside the image).
template <int N, int STRIDE, bool MASKED>
CS_DEVINL void bnrelu_vectors(uint8_t* p, const BnCoef8& k, const uint32_t* mask) {
  uint4 v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = *reinterpret_cast<const uint4*>(p + i * STRIDE);
  uint64_t x[N][4];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    x[i][0] = bf16x2_to_f32x2(v[i].x); x[i][1] = bf16x2_to_f32x2(v[i].y);
    x[i][2] = bf16x2_to_f32x2(v[i].z); x[i][3] = bf16x2_to_f32x2(v[i].w);
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) x[i][j] = fma_f32x2(x[i][j], k.sc[j], k.sh[j]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    uint4 o = make_uint4(relu_pack_bf16x2(x[i][0]), relu_pack_bf16x2(x[i][1]), relu_pack_bf16x2(x[i][2]), relu_pack_bf16x2(x[i][3]));
    if (MASKED) { o.x &= mask[i]; o.y &= mask[i]; o.z &= mask[i]; o.w &= mask[i]; }
    *reinterpret_cast<uint4*>(p + i * STRIDE) = o;
  }
}
// NROWS rows swept by NT threads, BATCH vectors per thread at a time; trip counts are compile-time constants (no branch
// between vectors); the NROWS % (NT / 8) rows left over are done by the first threads.
template <int PW, int NT, int NROWS, int BATCH>
CS_DEVINL void bnrelu_patch(uint8_t* slot, int t, const float* __restrict__ sc_g, const float* __restrict__ sh_g,
                            int h_first, int w_first, int H, int W, bool tile_valid) {
  static_assert(NT % 64 == 0, "rows per sweep must be a multiple of 8 (constant swizzle phase per thread)");
  constexpr int RS = NT / 8, NFULL = NROWS / RS, TAIL = NROWS - NFULL * RS;
  const int c = t & 7, r0 = t >> 3;
  BnCoef8 k;
  {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(sc_g + c * 8)), s1 = __ldg(reinterpret_cast<const float4*>(sc_g + c * 8 + 4));
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(sh_g + c * 8)), h1 = __ldg(reinterpret_cast<const float4*>(sh_g + c * 8 + 4));
    k.sc[0] = f32x2(s0.x, s0.y); k.sc[1] = f32x2(s0.z, s0.w); k.sc[2] = f32x2(s1.x, s1.y); k.sc[3] = f32x2(s1.z, s1.w);
    k.sh[0] = f32x2(h0.x, h0.y); k.sh[1] = f32x2(h0.z, h0.w); k.sh[2] = f32x2(h1.x, h1.y); k.sh[3] = f32x2(h1.z, h1.w);
  }
  uint8_t* p = slot + r0 * 128 + ((c ^ (r0 & 7)) << 4);
  const bool interior = tile_valid && h_first >= 0 && h_first + NROWS / PW <= H && w_first >= 0 && w_first + PW <= W;
  auto inside = [&](int r) -> uint32_t {                        // all ones if pixel row r lies inside the image
    const int ph = PW == 8 ? (r >> 3) : (r * 205) >> 11;        // r / 10 for r < 1029
    const int pw = r - ph * PW;
    return (tile_valid && (unsigned)(h_first + ph) < (unsigned)H && (unsigned)(w_first + pw) < (unsigned)W) ? 0xffffffffu : 0u;
  };
  if (interior) {
#pragma unroll
    for (int b0 = 0; b0 < NFULL; b0 += BATCH) {
      if (NFULL - b0 >= BATCH) bnrelu_vectors<BATCH, RS * 128, false>(p + b0 * RS * 128, k, nullptr);
      else bnrelu_vectors<(NFULL % BATCH ? NFULL % BATCH : BATCH), RS * 128, false>(p + b0 * RS * 128, k, nullptr);
    }
    if (TAIL && r0 < TAIL) bnrelu_vectors<1, RS * 128, false>(p + NFULL * RS * 128, k, nullptr);
  } else {
#pragma unroll
    for (int b0 = 0; b0 < NFULL; b0 += BATCH) {
      uint32_t m[BATCH];
#pragma unroll
      for (int i = 0; i < BATCH; ++i) m[i] = inside(r0 + (b0 + i) * RS);
      if (NFULL - b0 >= BATCH) bnrelu_vectors<BATCH, RS * 128, true>(p + b0 * RS * 128, k, m);
      else bnrelu_vectors<(NFULL % BATCH ? NFULL % BATCH : BATCH), RS * 128, true>(p + b0 * RS * 128, k, m);
    }
    if (TAIL && r0 < TAIL) {
      const uint32_t m = inside(r0 + NFULL * RS);
      bnrelu_vectors<1, RS * 128, true>(p + NFULL * RS * 128, k, &m);
    }
  }
}

// =============================================================================================
//                                   pixel-major GEMM
// =============================================================================================
// The pixel GEMMs are CTA-pair kernels: two CTAs (a cluster on one TPC) run ONE tcgen05.mma.cta_group::2 of M = 256
// pixels; each CTA stages its own 128-pixel A patches and HALF of the BLOCK_N weight rows, which halves the weight traffic
// from L2 and the B-operand shared-memory reads per CTA; each CTA's TMEM holds the accumulator rows of its own pixels.
// (Round 1's single-CTA kernel, kept then behind CARTSEG_PAIR=0, was removed in round 2: no test exercised it.)
// Epilogue of the CTA-pair pixel GEMMs (EG groups of 128 threads, warps 2..): TMEM -> registers -> (affine / ReLU) -> bf16 ->
// 128B-swizzled staging tile -> TMA store; train-mode BN statistics (sum, sum of squares) from the staged bf16 tile.
// Shared by pix_gemm2_kernel and conv3_gemm_kernel.
// EVALX: the eval-only extras (fused 1x1 head, 2x2 max-pooled copy) are compiled in.  They are a template parameter, not
// a run-time branch: carried by the training instantiations they cost 6-9 % of the Cout = 64 / conv-transpose kernels
// (registers and code in the loop that paces them; measured same-box, 16.2 -> 16.4 ms per K2 step).
template <int BLOCK_N, int EG, int BPG, bool EVALX, class Decode>
CS_DEVINL void pix_pair_epilogue(const PixGemmParams& p, uint8_t* stage_base, float* vec, float* red_base, uint64_t* tmem_full,
                                 uint64_t* tmem_empty, uint32_t tmem_base, int warp, int lane, int first_unit, int unit_stride,
                                 int num_units, bool want_stats, Decode&& decode) {
  const int eg = (warp - 2) >> 2;            // epilogue group: owns accumulator buffer `eg` when EG == 2
  const int et = threadIdx.x - 64 - eg * 128;
  const int sub = warp & 3;                  // TMEM sub-partition this warp may read
  const int ewarp = (warp - 2) & 3;
  const int bar0 = 3 * eg;                   // named barriers 1..3 (group 0), 4..6 (group 1)
  float* gvec = vec + eg * 2048;             // this group's statistics accumulators
  const int row = sub * 32 + lane;           // pixel row of the tile held by this thread
  float* red = red_base + eg * 512;
  int acc = 0, acc_phase = 0, unit_no = 0;
  uint32_t buf_ctr = 0;
  const bool affine = !want_stats && (p.scale != nullptr || p.shift != nullptr || p.relu);
  for (int u = first_unit; u < num_units; u += unit_stride, ++unit_no) {
    if (EG == 2) {                           // alternate units between the groups; acc buffer == group
      if ((unit_no & 1) != eg) continue;
      acc = eg;
      acc_phase = (unit_no >> 1) & 1;
    }
    int nb, b, w0, h0;
    const bool valid = decode(u, nb, b, w0, h0);

    mbar_wait(&tmem_full[acc], acc_phase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
    for (int cb = 0; cb < BLOCK_N / 64; ++cb) {
      // an n-block may span several output maps (conv-transpose: one map per (i,j) of the 2x2 kernel)
      const int col0 = nb * BLOCK_N + cb * 64;
      const int omap = col0 / p.cols_per_map;
      const int nin = col0 - omap * p.cols_per_map;     // first channel of this 64-wide chunk inside its map
      uint8_t* sbuf = stage_base + (eg * BPG + (int)(buf_ctr % BPG)) * kStageBytes;
      ++buf_ctr;
      if (et == 0) tma_store_wait_read<BPG - 1>();     // the store that last used this buffer has drained
      bar_sync(1 + bar0, 128);
      uint32_t v[64];
      tmem_ld32(taddr + cb * 64, v);
      tmem_ld32(taddr + cb * 64 + 32, v + 32);
      tmem_ld_wait();
      if (cb == BLOCK_N / 64 - 1) {                    // accumulator fully read: hand it back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(mapa_cluster(smem_u32(&tmem_empty[acc]), 0));
        }
      }
      uint32_t packed[32];
      if (affine) {
        const float* sc = vec + nin;
        const float* sh = vec + 1024 + nin;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float a = __uint_as_float(v[2 * i]) * sc[2 * i] + sh[2 * i];
          float c = __uint_as_float(v[2 * i + 1]) * sc[2 * i + 1] + sh[2 * i + 1];
          if (p.relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
          packed[i] = pack_bf16x2(a, c);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) packed[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      }
      if (EVALX && BLOCK_N == 64 && EG == 2 && p.head_logits != nullptr) {
        // fused 1x1 head (eval): this thread holds all 64 channels of its pixel; the activation tile is not stored
        const float* hw = vec + 2048;                      // head weights (group 1's statistics slots: unused in eval)
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          acc = fmaf(bf16_lo(packed[i]), hw[2 * i], acc);
          acc = fmaf(bf16_hi(packed[i]), hw[2 * i + 1], acc);
        }
        const int ph = h0 + (row >> 3), pw = w0 + (row & 7);
        if (valid && ph < p.H && pw < p.W) p.head_logits[((size_t)b * p.H + ph) * p.W + pw] = acc + hw[64];
        continue;
      }
      {
        uint8_t* rowp = sbuf + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 q = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) = q;
        }
      }
      fence_proxy_async();
      bar_sync(2 + bar0, 128);
      if (et == 0 && valid) {
        tma_store_4d(&p.tmapO[omap], sbuf, p.o_chan0 + nin, w0, h0, b);
        tma_store_commit();
      }
      if (EVALX && p.pool_out != nullptr && valid) {
        // 2x2 max-pool of the staged 8 x 16 pixel tile: 32 pooled pixels x 8 sixteen-byte chunks, two per thread
        uint8_t* pool = static_cast<uint8_t*>(p.pool_out);
        const int H2 = p.H >> 1, W2 = p.W >> 1;
#pragma unroll
        for (int q = et; q < 256; q += 128) {
          const int pp = q >> 3, cj = q & 7, ph = pp >> 2, pw = pp & 3;
          uint4 m = make_uint4(0u, 0u, 0u, 0u);             // post-ReLU values are >= 0
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            const int r = (2 * ph + (d >> 1)) * 8 + 2 * pw + (d & 1);
            const uint4 t4 = *reinterpret_cast<const uint4*>(sbuf + r * 128 + ((cj ^ (r & 7)) << 4));
            asm("max.bf16x2 %0, %0, %1;" : "+r"(m.x) : "r"(t4.x));
            asm("max.bf16x2 %0, %0, %1;" : "+r"(m.y) : "r"(t4.y));
            asm("max.bf16x2 %0, %0, %1;" : "+r"(m.z) : "r"(t4.z));
            asm("max.bf16x2 %0, %0, %1;" : "+r"(m.w) : "r"(t4.w));
          }
          const int oh = (h0 >> 1) + ph, ow = (w0 >> 1) + pw;
          if (oh < H2 && ow < W2) {                          // deep levels have ragged tiles (e.g. 12 x 12, 28 x 28 pixels)
            const size_t pix = ((size_t)b * H2 + oh) * W2 + ow;
            *reinterpret_cast<uint4*>(pool + (pix * p.Ntot + col0 + cj * 8) * 2) = m;
          }
        }
      }
      if (want_stats) {
        // Column sums over the staged bf16 tile: this warp covers rows [32*sub, 32*sub+32), lane
        // covers the channel pair (2*lane, 2*lane+1).  Tile rows that lie outside the image hold
        // values the store clips away; they are masked out here.
        // Packed fp32 arithmetic (add / fma .f32x2) on the (even, odd) channel pair, four independent accumulator pairs,
        // and no per-row branch when the tile lies inside the image: ncu showed this loop — 32 dependent, branchy rows of
        // 10 instructions — to be a third of the epilogue warps' time in the Cout = 64 kernels, which are paced by it.
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(sbuf);
        const int chunk = lane >> 2, word = lane & 3;
        uint64_t s2[4], q2[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { s2[a] = 0ull; q2[a] = 0ull; }
        const uint64_t one2 = f32x2(1.f, 1.f);
        if (valid) {
          const bool full = w0 + 8 <= p.W && h0 + 16 <= p.H;
          if (full) {
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
              const int r2 = sub * 32 + rr;
              const uint64_t x2 = bf16x2_to_f32x2(s32[r2 * 32 + (((chunk ^ (rr & 7)) << 2) | word)]);   // (32*sub + rr) & 7 == rr & 7
              s2[rr & 3] = fma_f32x2(x2, one2, s2[rr & 3]);
              q2[rr & 3] = fma_f32x2(x2, x2, q2[rr & 3]);
            }
          } else {
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
              const int r2 = sub * 32 + rr;
              const bool in = w0 + (r2 & 7) < p.W && h0 + (r2 >> 3) < p.H;
              const uint32_t w = s32[r2 * 32 + (((chunk ^ (rr & 7)) << 2) | word)];
              const uint64_t x2 = bf16x2_to_f32x2(in ? w : 0u);
              s2[rr & 3] = fma_f32x2(x2, one2, s2[rr & 3]);
              q2[rr & 3] = fma_f32x2(x2, x2, q2[rr & 3]);
            }
          }
        }
        float s0, s1, q0, q1;
        {
          const uint64_t sa = fma_f32x2(s2[1], one2, s2[0]), sb = fma_f32x2(s2[3], one2, s2[2]);
          const uint64_t qa = fma_f32x2(q2[1], one2, q2[0]), qb = fma_f32x2(q2[3], one2, q2[2]);
          const uint64_t st = fma_f32x2(sb, one2, sa), qt = fma_f32x2(qb, one2, qa);
          asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(st));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(qt));
        }
        float* rw = red + (ewarp * 64 + 2 * lane) * 2;
        rw[0] = s0; rw[1] = q0; rw[2] = s1; rw[3] = q1;
        bar_sync(3 + bar0, 128);
        const int c = et & 63, which = et >> 6;
        float tot = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) tot += red[(w4 * 64 + c) * 2 + which];
        gvec[which * 1024 + nin + c] += tot;
      }
    }
    if (EG == 1) {
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  if (et == 0) tma_store_wait_all<0>();
  if (want_stats) {
    bar_sync(1 + bar0, 128);
    const int nvec = p.cols_per_map;
    for (int i = et; i < nvec; i += 128) {
      const float s = gvec[i], q = gvec[1024 + i];
      if (q != 0.f) {
        atomicAdd(&p.stat_sum[i], (double)s);
        atomicAdd(&p.stat_sq[i], (double)q);
      }
    }
  }
}

// Pair-only kernel with unified pipeline stages (A patch + its R weight tiles per stage, one mbarrier wait and one
// commit per 4*R MMAs): ncu showed the single MMA-issuing warp to be the bottleneck of the finer-grained ring.
// EG = number of 4-warp epilogue groups: with EG = 2 the groups take alternate work units (group e owns TMEM
// accumulator buffer e), which doubles the epilogue throughput of the small-N kernels whose BN-statistics epilogue
// was slower than their MMAs (ncu: the MMA issuer waited on tmem_empty).
template <int BLOCK_N, int S, int NSTG, int EG>
struct Pix2Layout {
  static constexpr int kBRows = BLOCK_N / 2;
  static constexpr int kBSlot = kBRows * 128;
  static constexpr int kStageSz = kASlotBytes + 3 * kBSlot;
  static constexpr int kStage = S * kStageSz;                // epilogue staging
  static constexpr int kVec = kStage + NSTG * kStageBytes;   // EG x (2 x 1024 floats)
  static constexpr int kRed = kVec + EG * 2 * 1024 * 4;      // EG x (4 x 64 x 2 floats)
  static constexpr int kBar = kRed + EG * 4 * 64 * 2 * 4;
  static constexpr int kThreadsTotal = 64 + 128 * EG;
  static constexpr int kNumBar = 2 * S + 4;
  static constexpr int kTmemPtr = kBar + kNumBar * 8;
  static constexpr int kTotal = kTmemPtr + 16;
  static constexpr int kDyn = kTotal + 1024;                 // slack for manual 1024-B alignment
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
};

template <int BLOCK_N, int S, int NSTG, int EG>
__global__ void __launch_bounds__(64 + 128 * EG, 1) pix_gemm2_kernel(const __grid_constant__ PixGemmParams p) {
  using L = Pix2Layout<BLOCK_N, S, NSTG, EG>;
  constexpr bool PAIR = true;
  static_assert(NSTG % EG == 0, "every epilogue group owns NSTG / EG staging buffers");
  constexpr int BPG = NSTG / EG;                             // staging buffers per epilogue group (used round-robin)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* full = bars;
  uint64_t* empty = full + S;
  uint64_t* tmem_full = empty + S;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtr);
  float* vec = reinterpret_cast<float*>(smem + L::kVec);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;       // CTA rank inside the pair
  const bool leader = rank == 0;
  const int m_tiles = p.tiles_w * p.tiles_h * p.batch;
  // work unit: PAIR ? (two consecutive m-tiles) x one n-block : one m-tile x one n-block
  const int m_units = PAIR ? (m_tiles + 1) / 2 : m_tiles;
  const int num_units = m_units * p.n_blocks;
  const int first_unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const bool want_stats = p.stat_sum != nullptr;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], PAIR ? 8 : 4); }
    fence_barrier_init();
    for (int g = 0; g < p.G; ++g) tma_prefetch_desc(&p.tmapA[p.a_map[g]]);
    tma_prefetch_desc(&p.tmapB);
    tma_prefetch_desc(&p.tmapO[0]);
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair<L::kTmemCols>(tmem_ptr);
    else tmem_alloc<L::kTmemCols>(tmem_ptr);
  }
  {
    const int nvec = p.cols_per_map;                         // <= 1024
    for (int i = threadIdx.x; i < 1024; i += L::kThreadsTotal) {
      float a = 0.f, b = 0.f;
      if (!want_stats && i < nvec) {
        a = p.scale ? p.scale[i] : 1.f;
        b = p.shift ? p.shift[i] : 0.f;
      }
      vec[i] = a;                                          // group 0's copy doubles as the scale / shift table
      vec[1024 + i] = b;
      if (EG == 2) { vec[2048 + i] = 0.f; vec[3072 + i] = 0.f; }
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync();                                  // both CTAs: barriers initialised, TMEM allocated
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int per_img = p.tiles_w * p.tiles_h;
  // Decodes work unit u into this CTA's tile; a CTA whose m-tile does not exist (odd tile count) gets batch
  // coordinate == batch: every TMA load is zero-filled and every store clipped away.
  auto decode = [&](int u, int& nb, int& b, int& w0, int& h0) -> bool {
    const int m_unit = u / p.n_blocks;
    nb = u - m_unit * p.n_blocks;
    const int m_tile = PAIR ? 2 * m_unit + (int)rank : m_unit;
    const bool valid = m_tile < m_tiles;
    b = valid ? m_tile / per_img : p.batch;
    const int rem = valid ? m_tile - b * per_img : 0;
    const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
    w0 = tw * 8;
    h0 = th * 16;
    return valid;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // One pipeline stage = the A patch of one (64-channel chunk, horizontal tap) + the weight tiles of its R
    // vertical taps, all landing on one mbarrier (the leader's, for both CTAs of the pair).
    int st = 0, ph = 0;
    const uint32_t a_bytes = (16 + p.R - 1) * kAtomBytes;
    const uint32_t stage_bytes = 2 * (a_bytes + p.R * L::kBSlot);
    for (int u = first_unit; u < num_units; u += unit_stride) {
      int nb, b, w0, h0;
      decode(u, nb, b, w0, h0);
      const int n0 = nb * BLOCK_N + (int)rank * L::kBRows;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        for (int g = 0; g < p.G; ++g) {
          mbar_wait(&empty[st], ph ^ 1);
          if (lane == 0) {
            uint8_t* base = smem + st * L::kStageSz;
            const uint32_t bar = mapa_cluster(smem_u32(&full[st]), 0);
            if (leader) mbar_arrive_expect_tx(&full[st], stage_bytes);
            tma_load_4d_pair(base, &p.tmapA[p.a_map[g]], bar, p.a_chan0 + kc * 64, w0 + p.a_dw[g], h0 + p.a_dh[g], b);
            for (int r = 0; r < p.R; ++r)
              tma_load_2d_pair(base + kASlotBytes + r * L::kBSlot, &p.tmapB, bar, kc * 64, (g * p.R + r) * p.Ntot + n0);
          }
          __syncwarp();
          if (++st == S) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      // Every lane runs this (warp-uniform) code; one elected lane issues each tcgen05 instruction.  The loop is
      // organised around the smem ring (stage index = compile-time constant of the unrolled body) so that every
      // descriptor is "uniform base + immediate": the single issuing warp is the critical path of this kernel.
      constexpr uint32_t idesc = make_idesc(256, BLOCK_N, 0, 0);
      const uint32_t lbo_lo = (16u >> 4) << 16;                             // LBO field lives in the low word
      const uint32_t s_base = (smem_u32(smem) >> 4) | lbo_lo;
      const int per_unit = p.kchunks * p.G;
      int ph = 0, acc = 0, acc_phase = 0, kg = 0, u = first_unit;
      uint32_t accumulate = 0;
      bool more = u < num_units;
      while (more) {
#pragma unroll
        for (int st = 0; st < S; ++st) {
          if (more) {
            if (kg == 0) {
              mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
              accumulate = 0;
            }
            mbar_wait(&full[st], ph);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            const uint32_t a_lo = s_base + (uint32_t)(st * (L::kStageSz >> 4));
            const uint32_t b_lo = a_lo + (kASlotBytes >> 4);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              if (r < p.R) {
                // four K-steps of 16 channels: +32 bytes = +2 in the descriptor start field per step
                umma_bf16_steps_warp<true, 4, 2>(d_tmem, a_lo + (uint32_t)(r * (kAtomBytes >> 4)),
                                                 b_lo + (uint32_t)(r * (L::kBSlot >> 4)), idesc, accumulate);
                accumulate = 1;
              }
            }
            umma_commit_warp<true>(&empty[st]);
            if (++kg == per_unit) {
              umma_commit_warp<true>(&tmem_full[acc]);
              kg = 0;
              acc ^= 1;
              if (acc == 0) acc_phase ^= 1;
              u += unit_stride;
              more = u < num_units;
            }
          }
        }
        ph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (EG groups of 128 threads)
    pix_pair_epilogue<BLOCK_N, EG, BPG, false>(p, smem + L::kStage, vec, reinterpret_cast<float*>(smem + L::kRed), tmem_full, tmem_empty,
                                               tmem_base, warp, lane, first_unit, unit_stride, num_units, want_stats, decode);
  }
  tc_fence_before();
  if (PAIR) cluster_sync();                                  // the peer's smem / TMEM / barriers stay alive until here
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<L::kTmemCols>(tmem_base);
    else tmem_dealloc<L::kTmemCols>(tmem_base);
  }
}

// =============================================================================================
//                       3x3 convolution with ONE activation patch per K-chunk
// =============================================================================================
// pix_gemm2_kernel loads an 8 x 18 pixel patch per (64-channel chunk, horizontal tap): every activation byte crosses
// L2 -> shared memory three times, and the kernel is paced by that traffic, not by the MMAs (ncu, round 1: 9.4 TB/s of
// l1tex<-xbar reads at N = 256; at N = 64 the loads per MMA-cycle are 3.6x higher still: 0.45 of the tensor peak).
// Here ONE (8+2) x (16+2) patch per 64-channel chunk serves all nine taps: tap (kh, kw) of the tile is the same patch
// read from row offset (kh * PW + kw) with a row-group stride (SBO) of PW * 128 B — the 128B swizzle is a function of
// the shared-memory address, so the view stays consistent with what TMA wrote.  Weights: a ring of 3-tap groups
// (RESIDENT = false), or, for the Cout = 64 layers (n_blocks == 1, K <= 128), ALL taps resident in shared memory for the
// whole kernel (73 KB): per work unit only the 23 KB patch moves.
template <int BLOCK_N, int PW, int SA, int SB, int NSTG, int EG, bool RESIDENT, bool XF = false>
struct Conv3Layout {
  static constexpr int kAPatch = 18 * PW * 128;
  static constexpr int kASlot = (kAPatch + 1023) & ~1023;
  static constexpr int kBRows = BLOCK_N / 2;
  static constexpr int kBTile = kBRows * 128;
  static constexpr int kBGroup = 3 * kBTile;
  static constexpr int kMaxResidentChunks = 2;
  static constexpr int kBBytes = RESIDENT ? kMaxResidentChunks * 9 * kBTile : SB * kBGroup;
  static constexpr int kNumB = RESIDENT ? 1 : SB;
  static constexpr int kA = 0;
  static constexpr int kB = kA + SA * kASlot;
  static constexpr int kStage = kB + kBBytes;
  static constexpr int kVec = kStage + NSTG * kStageBytes;   // EG x (2 x 1024 floats)
  static constexpr int kRed = kVec + EG * 2 * 1024 * 4;      // EG x (4 x 64 x 2 floats)
  static constexpr int kBar = kRed + EG * 4 * 64 * 2 * 4;
  static constexpr int kXfWarps = XF ? (EG == 2 ? 4 : 2) : 0;                // transform warps (BatchNorm + ReLU of every A patch)
  static constexpr int kThreadsTotal = 64 + 128 * EG + 32 * kXfWarps;
  static constexpr int kNumBar = 2 * SA + 2 * kNumB + 4 + (XF ? SA : 0);
  static constexpr int kTmemPtr = kBar + kNumBar * 8;
  static constexpr int kTotal = kTmemPtr + 16;
  static constexpr int kDyn = kTotal + 1024;                 // slack for manual 1024-B alignment
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static_assert(RESIDENT || (3 * SA) % SB == 0, "the weight-ring slot of (A slot, tap group) must be a compile-time constant");
};

template <int BLOCK_N, int PW, int SA, int SB, int NSTG, int EG, bool RESIDENT, bool XF, bool EVALX>
__global__ void __launch_bounds__(64 + 128 * EG + (XF ? (EG == 2 ? 128 : 64) : 0), 1) conv3_gemm_kernel(const __grid_constant__ PixGemmParams p) {
  using L = Conv3Layout<BLOCK_N, PW, SA, SB, NSTG, EG, RESIDENT, XF>;
  static_assert(NSTG % EG == 0, "every epilogue group owns NSTG / EG staging buffers");
  constexpr int BPG = NSTG / EG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + SA;
  uint64_t* fullB = emptyA + SA;
  uint64_t* emptyB = fullB + L::kNumB;
  uint64_t* tmem_full = emptyB + L::kNumB;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* landA = tmem_empty + 2;                          // XF: this CTA's patch has landed (local), fullA = transformed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtr);
  float* vec = reinterpret_cast<float*>(smem + L::kVec);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_tiles = p.tiles_w * p.tiles_h * p.batch;
  const int m_units = (m_tiles + 1) / 2;                     // two consecutive m-tiles (one per CTA of the pair)
  const int num_units = m_units * p.n_blocks;
  const int first_unit = (int)(blockIdx.x >> 1);
  const int unit_stride = (int)(gridDim.x >> 1);
  const bool want_stats = p.stat_sum != nullptr;

  if (threadIdx.x == 0) {
    // XF: a patch is ready for the MMAs when the transform warps of BOTH CTAs have arrived on the leader's barrier
    for (int i = 0; i < SA; ++i) { mbar_init(&fullA[i], XF ? 2 * L::kXfWarps : 1); mbar_init(&emptyA[i], 1); }
    if (XF) for (int i = 0; i < SA; ++i) mbar_init(&landA[i], 1);
    for (int i = 0; i < L::kNumB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
    fence_barrier_init();
    tma_prefetch_desc(&p.tmapA3);
    tma_prefetch_desc(&p.tmapB);
    tma_prefetch_desc(&p.tmapO[0]);
  }
  if (warp == 2) tmem_alloc_pair<L::kTmemCols>(tmem_ptr);
  {
    const int nvec = p.cols_per_map;                         // <= 1024
    for (int i = threadIdx.x; i < 1024; i += L::kThreadsTotal) {
      float a = 0.f, b = 0.f;
      if (!want_stats && i < nvec) {
        a = p.scale ? p.scale[i] : 1.f;
        b = p.shift ? p.shift[i] : 0.f;
      }
      vec[i] = a;                                            // group 0's copy doubles as the scale / shift table
      vec[1024 + i] = b;
      if (EG == 2) { vec[2048 + i] = 0.f; vec[3072 + i] = 0.f; }
    }
    if (EVALX && BLOCK_N == 64 && EG == 2 && p.head_logits != nullptr) {   // eval-mode fused head: weights [0, 64), bias at [64]
      __syncthreads();
      if (threadIdx.x < 64) vec[2048 + threadIdx.x] = p.head_w[threadIdx.x];
      if (threadIdx.x == 64) vec[2048 + 64] = p.head_b ? p.head_b[0] : 0.f;
    }
  }
  tc_fence_before();
  cluster_sync();                                            // both CTAs: barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int per_img = p.tiles_w * p.tiles_h;
  auto decode = [&](int u, int& nb, int& b, int& w0, int& h0) -> bool {
    const int m_unit = u / p.n_blocks;
    nb = u - m_unit * p.n_blocks;
    const int m_tile = 2 * m_unit + (int)rank;
    const bool valid = m_tile < m_tiles;
    b = valid ? m_tile / per_img : p.batch;                  // batch coordinate == batch: loads zero-filled, stores clipped
    const int rem = valid ? m_tile - b * per_img : 0;
    const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
    w0 = tw * 8;
    h0 = th * 16;
    return valid;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The activation patch of K-chunk q+1 is requested right after the first weight group of chunk q, so that the
    // weight ring never holds the patches back by more than one group.
    int a_u = first_unit, a_kc = 0, sa = 0, pa = 0;
    bool a_more = a_u < num_units;
    auto issue_A = [&]() {
      int nb, b, w0, h0;
      decode(a_u, nb, b, w0, h0);
      mbar_wait(&emptyA[sa], pa ^ 1);
      if (lane == 0) {
        if (XF) {                                              // lands on this CTA's own barrier: its transform warps wait there
          mbar_arrive_expect_tx(&landA[sa], L::kAPatch);
          tma_load_4d(smem + L::kA + sa * L::kASlot, &p.tmapA3, &landA[sa], p.a_chan0 + a_kc * 64, w0 - 1, h0 - 1, b);
        } else {
          if (leader) mbar_arrive_expect_tx(&fullA[sa], 2 * L::kAPatch);
          tma_load_4d_pair(smem + L::kA + sa * L::kASlot, &p.tmapA3, mapa_cluster(smem_u32(&fullA[sa]), 0),
                           p.a_chan0 + a_kc * 64, w0 - 1, h0 - 1, b);
        }
      }
      __syncwarp();
      if (++sa == SA) { sa = 0; pa ^= 1; }
      if (++a_kc == p.kchunks) { a_kc = 0; a_u += unit_stride; a_more = a_u < num_units; }
    };
    if (RESIDENT) {
      if (lane == 0 && first_unit < num_units) {
        const uint32_t bar = mapa_cluster(smem_u32(&fullB[0]), 0);
        if (leader) mbar_arrive_expect_tx(&fullB[0], 2 * p.kchunks * 9 * L::kBTile);
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int t = 0; t < 9; ++t)
            tma_load_2d_pair(smem + L::kB + (kc * 9 + t) * L::kBTile, &p.tmapB, bar, kc * 64, t * p.Ntot + (int)rank * L::kBRows);
      }
      __syncwarp();
      while (a_more) issue_A();
    } else {
      int sb = 0, pb = 0;
      if (a_more) issue_A();
      for (int u = first_unit; u < num_units; u += unit_stride) {
        const int nb = u % p.n_blocks;
        const int n0 = nb * BLOCK_N + (int)rank * L::kBRows;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int g = 0; g < 3; ++g) {
            mbar_wait(&emptyB[sb], pb ^ 1);
            if (lane == 0) {
              const uint32_t bar = mapa_cluster(smem_u32(&fullB[sb]), 0);
              if (leader) mbar_arrive_expect_tx(&fullB[sb], 2 * L::kBGroup);
              for (int r = 0; r < 3; ++r)
                tma_load_2d_pair(smem + L::kB + sb * L::kBGroup + r * L::kBTile, &p.tmapB, bar, kc * 64,
                                 (g * 3 + r) * p.Ntot + n0);
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; pb ^= 1; }
            if (g == 0 && a_more) issue_A();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc(256, BLOCK_N, 0, 0);
      const uint32_t lbo_lo = (16u >> 4) << 16;                             // LBO field lives in the low word
      const uint32_t s_base = (smem_u32(smem) >> 4) | lbo_lo;
      // A descriptor high word: SBO = one image row of the patch (PW pixels x 128 B), version 1, 128B swizzle
      // (the descriptor's base-offset field stays 0: measured on B200, the swizzle is applied to the absolute shared-
      // memory address, so a view that starts k rows into a 1024-byte atom needs no correction — setting the field to
      // the row phase, as the PTX text for non-aligned matrices suggests, gives wrong results)
      constexpr uint32_t a_hi = ((uint32_t)(PW * 128) >> 4) | (1u << 14) | (2u << 29);
      int kc = 0, acc = 0, acc_phase = 0, phA = 0, u = first_unit;
      uint32_t gq = 0;                                                      // weight groups consumed so far
      uint32_t accumulate = 0;
      bool more = u < num_units;
      if (RESIDENT && more) { mbar_wait(&fullB[0], 0); tc_fence_after(); }
      while (more) {
#pragma unroll
        for (int st = 0; st < SA; ++st) {
          if (more) {
            if (kc == 0) {
              mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
              accumulate = 0;
            }
            mbar_wait(&fullA[st], phA);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            const uint32_t a_slot = s_base + (uint32_t)((L::kA + st * L::kASlot) >> 4);
#pragma unroll
            for (int g = 0; g < 3; ++g) {
              uint32_t b_lo;
              constexpr int kSlotMod = RESIDENT ? 1 : SB;
              const int slot = RESIDENT ? 0 : (3 * st + g) % kSlotMod;
              if (RESIDENT) {
                b_lo = s_base + (uint32_t)(L::kB >> 4) + (uint32_t)((kc * 9 + g * 3) * (L::kBTile >> 4));
              } else {
                mbar_wait(&fullB[slot], (gq / (uint32_t)kSlotMod) & 1u);
                tc_fence_after();
                b_lo = s_base + (uint32_t)((L::kB + slot * L::kBGroup) >> 4);
              }
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const uint32_t row0 = (uint32_t)(r * PW + g);               // first patch row of this tap's view
                // four K-steps of 16 channels: +32 bytes = +2 in the descriptor start field per step
                umma_bf16_steps4_warp_hi<true>(d_tmem, a_slot + row0 * 8u, a_hi, b_lo + (uint32_t)(r * (L::kBTile >> 4)), idesc,
                                               accumulate);
                accumulate = 1;
              }
              if (!RESIDENT) umma_commit_warp<true>(&emptyB[slot]);
              ++gq;
            }
            umma_commit_warp<true>(&emptyA[st]);
            if (++kc == p.kchunks) {
              umma_commit_warp<true>(&tmem_full[acc]);
              kc = 0;
              acc ^= 1;
              if (acc == 0) acc_phase ^= 1;
              u += unit_stride;
              more = u < num_units;
            }
          }
        }
        phA ^= 1;
      }
    }
  } else if (XF && warp >= 2 + 4 * EG) {
    // ------------------------------------------------------------------ transform warps (both CTAs): BN + ReLU of each patch
    const int tt = threadIdx.x - (2 + 4 * EG) * 32;          // 0..63
    int sa = 0, pa = 0;
    for (int u = first_unit; u < num_units; u += unit_stride) {
      int nb, b, w0, h0;
      const bool valid = decode(u, nb, b, w0, h0);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&landA[sa], pa);
        bnrelu_patch<PW, (XF ? 32 * L::kXfWarps : 64), 18 * PW, 6>(smem + L::kA + sa * L::kASlot, tt, p.in_scale + kc * 64, p.in_shift + kc * 64,
                                           h0 - 1, w0 - 1, p.H, p.W, valid);
        fence_proxy_async();                                  // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&fullA[sa]), 0));
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
    }
  } else {
    pix_pair_epilogue<BLOCK_N, EG, BPG, EVALX>(p, smem + L::kStage, vec, reinterpret_cast<float*>(smem + L::kRed), tmem_full, tmem_empty,
                                               tmem_base, warp, lane, first_unit, unit_stride, num_units, want_stats, decode);
  }
  tc_fence_before();
  cluster_sync();                                            // the peer's smem / TMEM / barriers stay alive until here
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<L::kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N, int PW, int SA, int SB, int NSTG, int EG, bool RESIDENT, bool XF, bool EVALX = false>
static cudaError_t launch_conv3(const PixGemmParams& p, int num_sms, cudaStream_t stream) {
  using L = Conv3Layout<BLOCK_N, PW, SA, SB, NSTG, EG, RESIDENT, XF>;
  static_assert(L::kDyn <= 232448, "shared memory budget exceeded");
  auto kern = conv3_gemm_kernel<BLOCK_N, PW, SA, SB, NSTG, EG, RESIDENT, XF, EVALX>;
  static std::atomic<unsigned long long> attr_done{0};
  {
    cudaError_t ae = ensure_dynamic_smem(kern, L::kDyn, attr_done);
    if (ae != cudaSuccess) return ae;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * p.batch;
  const int units = ((m_tiles + 1) / 2) * p.n_blocks;
  if (units <= 0) return cudaSuccess;
  const int clusters = units < num_sms / 2 ? units : num_sms / 2;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(2 * clusters);
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.blockDim = dim3(L::kThreadsTotal);
  cfg.dynamicSmemBytes = L::kDyn;
  cfg.stream = stream;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return e;
  return launched();
}

// =============================================================================================
//                                   stem: first convolution (Cin <= 7)
// =============================================================================================
// K = 9 * Cin <= 63: one 64-wide K chunk.  Round 1 wrote the im2col matrix [pixels][64] bf16 to HBM (347 MB for a 38 MB
// image at K2) and read it back through the pointwise GEMM: 0.16 + 0.19 ms.  Here each 128-pixel A tile is built
// directly in (128B-swizzled) shared memory: warp 0 brings the fp32 patch of the tile — (8+8) x (16+2) pixels of every
// input channel, through a 4-D TMA map over the NCHW image whose out-of-bounds zero fill is the convolution padding —
// into a small ring, two builder warps turn it into im2col rows (27 shared-memory reads per pixel instead of 27 scattered
// global loads: the first version of this kernel, which read the image with __ldg, took 0.55 ms) and hand the tile to
// the MMA warp through an mbarrier; the 64 x 64 weight tile is resident; the epilogue (bf16 store + BN statistics) is the
// shared pix_pair_epilogue.  The kernel is paced by its 411 MB output.
// im2col row: k = (kh*3 + kw) * Cin + c, zero beyond 9 * Cin.  The 16-byte chunks past ceil(9*Cin / 8) are zeroed once
// per kernel and never rewritten (their swizzled position depends on the tile row only).
// 10 columns (w0-1 .. w0+8) are needed; the box starts at w0-4 and is 16 wide so that the innermost TMA coordinate is a
// multiple of 16 bytes (a box starting at w0-1 — 12 columns — faults with "illegal instruction" on B200)
static constexpr int kStemPatchW = 16;
static constexpr int kStemPatchX0 = 3;                       // column of tap kw = 0 for tile pixel w = 0
static constexpr int kStemPatchH = 18;

template <int SA, int SX, int EG>
struct StemLayout {
  static constexpr int kASlot = 128 * 128;                   // 128 pixels x 64 k (bf16)
  static constexpr int kBTile = 32 * 128;                    // this CTA's half of the 64 output channels
  static constexpr int kXSlot = ((kStemPatchW * kStemPatchH * 7 * 4) + 1023) & ~1023;  // fp32 patch, Cin <= 7; keeps kStage 1024-aligned
  static constexpr int kA = 0;
  static constexpr int kB = kA + SA * kASlot;
  static constexpr int kX = kB + kBTile;
  static constexpr int kStage = kX + SX * kXSlot;
  static constexpr int kVec = kStage + EG * kStageBytes;
  static constexpr int kRed = kVec + EG * 2 * 1024 * 4;
  static constexpr int kBar = kRed + EG * 4 * 64 * 2 * 4;
  static constexpr int kBuilderWarps = 2;
  static constexpr int kThreadsTotal = 64 + 128 * EG + 32 * kBuilderWarps;
  static constexpr int kNumBar = 2 * SA + 2 * SX + 1 + 4;
  static constexpr int kTmemPtr = kBar + kNumBar * 8;
  static constexpr int kTotal = kTmemPtr + 16;
  static constexpr int kDyn = kTotal + 1024;
  static constexpr int kTmemCols = 128;                      // two 64-column accumulators
  static_assert(kStage % 1024 == 0 && kB % 1024 == 0, "128B-swizzled tiles must start on a 1024-byte boundary");
};

template <int SA, int SX, int EG>
__global__ void __launch_bounds__(64 + 128 * EG + 64, 1) stem_gemm_kernel(const __grid_constant__ PixGemmParams p) {
  using L = StemLayout<SA, SX, EG>;
  constexpr int BLOCK_N = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + SA;
  uint64_t* fullX = emptyA + SA;
  uint64_t* emptyX = fullX + SX;
  uint64_t* fullB = emptyX + SX;
  uint64_t* tmem_full = fullB + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtr);
  float* vec = reinterpret_cast<float*>(smem + L::kVec);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_tiles = p.tiles_w * p.tiles_h * p.batch;
  const int num_units = (m_tiles + 1) / 2;                   // n_blocks == 1
  const int first_unit = (int)(blockIdx.x >> 1);
  const int unit_stride = (int)(gridDim.x >> 1);
  const bool want_stats = p.stat_sum != nullptr;

  if (threadIdx.x == 0) {
    // a tile is complete when the builder warps of BOTH CTAs have arrived on the leader's barrier
    for (int i = 0; i < SA; ++i) { mbar_init(&fullA[i], 2 * L::kBuilderWarps); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < SX; ++i) { mbar_init(&fullX[i], 1); mbar_init(&emptyX[i], L::kBuilderWarps); }
    mbar_init(fullB, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
    fence_barrier_init();
    tma_prefetch_desc(&p.tmapB);
    tma_prefetch_desc(&p.tmapO[0]);
    tma_prefetch_desc(&p.tmapX);
  }
  if (warp == 2) tmem_alloc_pair<L::kTmemCols>(tmem_ptr);
  {
    const int nvec = p.cols_per_map;
    for (int i = threadIdx.x; i < 1024; i += L::kThreadsTotal) {
      float a = 0.f, b = 0.f;
      if (!want_stats && i < nvec) {
        a = p.scale ? p.scale[i] : 1.f;
        b = p.shift ? p.shift[i] : 0.f;
      }
      vec[i] = a;
      vec[1024 + i] = b;
      if (EG == 2) { vec[2048 + i] = 0.f; vec[3072 + i] = 0.f; }
    }
    // zero the A ring once: the builders only ever write the chunks that hold k < 9 * Cin
    uint4* a4 = reinterpret_cast<uint4*>(smem + L::kA);
    for (int i = threadIdx.x; i < SA * L::kASlot / 16; i += L::kThreadsTotal) a4[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int per_img = p.tiles_w * p.tiles_h;
  auto decode = [&](int u, int& nb, int& b, int& w0, int& h0) -> bool {
    nb = 0;
    const int m_tile = 2 * u + (int)rank;
    const bool valid = m_tile < m_tiles;
    b = valid ? m_tile / per_img : p.batch;
    const int rem = valid ? m_tile - b * per_img : 0;
    const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
    w0 = tw * 8;
    h0 = th * 16;
    return valid;
  };

  constexpr int kFirstBuilderWarp = 2 + 4 * EG;
  const int Cin = p.stem_cin;
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA: weights once, then one fp32 patch per tile
    if (lane == 0 && first_unit < num_units) {
      if (leader) mbar_arrive_expect_tx(fullB, 2 * L::kBTile);
      tma_load_2d_pair(smem + L::kB, &p.tmapB, mapa_cluster(smem_u32(fullB), 0), 0, (int)rank * 32);
    }
    __syncwarp();
    const uint32_t patch_bytes = (uint32_t)(kStemPatchW * kStemPatchH * 4 * Cin);
    int sx = 0, px = 0;
    for (int u = first_unit; u < num_units; u += unit_stride) {
      int nb, b, w0, h0;
      decode(u, nb, b, w0, h0);
      mbar_wait(&emptyX[sx], px ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&fullX[sx], patch_bytes);
        tma_load_4d(smem + L::kX + sx * L::kXSlot, &p.tmapX, &fullX[sx], w0 - 1 - kStemPatchX0, h0 - 1, 0, b);
      }
      __syncwarp();
      if (++sx == SX) { sx = 0; px ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && first_unit < num_units) {
      constexpr uint32_t idesc = make_idesc(256, BLOCK_N, 0, 0);
      const uint32_t lbo_lo = (16u >> 4) << 16;
      const uint32_t s_base = (smem_u32(smem) >> 4) | lbo_lo;
      const uint32_t b_lo = s_base + (uint32_t)(L::kB >> 4);
      mbar_wait(fullB, 0);
      tc_fence_after();
      int st = 0, ph = 0, acc = 0, acc_phase = 0;
      for (int u = first_unit; u < num_units; u += unit_stride) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        mbar_wait(&fullA[st], ph);
        tc_fence_after();
        umma_bf16_steps_warp<true, 4, 2>(tmem_base + acc * BLOCK_N, s_base + (uint32_t)((L::kA + st * L::kASlot) >> 4), b_lo, idesc, 0);
        umma_commit_warp<true>(&emptyA[st]);
        umma_commit_warp<true>(&tmem_full[acc]);
        if (++st == SA) { st = 0; ph ^= 1; }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= kFirstBuilderWarp) {
    // ------------------------------------------------------------------ A builders: 64 threads, two pixel rows each
    const int bt = threadIdx.x - kFirstBuilderWarp * 32;     // 0..63
    int st = 0, ph = 0, sx = 0, px = 0;
    for (int u = first_unit; u < num_units; u += unit_stride) {
      mbar_wait(&fullX[sx], px);
      mbar_wait(&emptyA[st], ph ^ 1);
      const float* xs = reinterpret_cast<const float*>(smem + L::kX + sx * L::kXSlot);
      uint8_t* slot = smem + L::kA + st * L::kASlot;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int row = bt + 64 * rr;
        const float* xr = xs + (row >> 3) * kStemPatchW + (row & 7) + kStemPatchX0;   // tap (0, 0) of channel 0
        uint8_t* rowp = slot + row * 128;
        if (Cin == 3) {
          float v[28];
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[tap * 3 + c] = xr[(c * kStemPatchH + tap / 3) * kStemPatchW + tap % 3];
          v[27] = 0.f;
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 14; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          pk[14] = 0u; pk[15] = 0u;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        } else {
          // generic 1..7 channels (not the benchmarked configuration): element-wise 2-byte stores
          const int K = 9 * Cin;
          for (int k = 0; k < K; ++k) {
            const int tap = k / Cin, c = k - tap * Cin;
            const float f = xr[(c * kStemPatchH + tap / 3) * kStemPatchW + tap % 3];
            *reinterpret_cast<__nv_bfloat16*>(rowp + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1))) = __float2bfloat16(f);
          }
        }
      }
      fence_proxy_async();                                    // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&emptyX[sx]);                             // the fp32 patch may be overwritten
        mbar_arrive_cluster(mapa_cluster(smem_u32(&fullA[st]), 0));
      }
      if (++st == SA) { st = 0; ph ^= 1; }
      if (++sx == SX) { sx = 0; px ^= 1; }
    }
  } else {
    pix_pair_epilogue<BLOCK_N, EG, 1, false>(p, smem + L::kStage, vec, reinterpret_cast<float*>(smem + L::kRed), tmem_full, tmem_empty,
                                             tmem_base, warp, lane, first_unit, unit_stride, num_units, want_stats, decode);
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<L::kTmemCols>(tmem_base);
  }
}

static cudaError_t launch_stem_gemm(const PixGemmParams& p, int num_sms, cudaStream_t stream) {
  constexpr int SA = 4, SX = 4, EG = 2;
  using L = StemLayout<SA, SX, EG>;
  static_assert(L::kDyn <= 232448, "shared memory budget exceeded");
  auto kern = stem_gemm_kernel<SA, SX, EG>;
  static std::atomic<unsigned long long> attr_done{0};
  {
    cudaError_t ae = ensure_dynamic_smem(kern, L::kDyn, attr_done);
    if (ae != cudaSuccess) return ae;
  }
  if (p.n_blocks != 1 || p.Ntot != 64 || p.kchunks != 1 || p.stem_cin < 1 || p.stem_cin > 7) return cudaErrorInvalidValue;
  const int m_tiles = p.tiles_w * p.tiles_h * p.batch;
  const int units = (m_tiles + 1) / 2;
  if (units <= 0) return cudaSuccess;
  const int clusters = units < num_sms / 2 ? units : num_sms / 2;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(2 * clusters);
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.blockDim = dim3(L::kThreadsTotal);
  cfg.dynamicSmemBytes = L::kDyn;
  cfg.stream = stream;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return e;
  return launched();
}

static cudaError_t launch_conv3_gemm(const PixGemmParams& p, int block_n, int num_sms, cudaStream_t stream) {
  if (p.G != 3 || p.R != 3 || !p.pair) return cudaErrorInvalidValue;
  const bool resident = block_n == 64 && p.n_blocks == 1 && p.kchunks <= 2;
  if (p.in_scale && p.in_shift) {                            // operand = raw conv output: BN + ReLU applied in shared memory
    switch (block_n) {
      case 64: return resident ? launch_conv3<64, 10, 4, 3, 2, 2, true, true>(p, num_sms, stream)
                               : launch_conv3<64, 10, 4, 3, 2, 2, false, true>(p, num_sms, stream);
      case 128: return launch_conv3<128, 10, 4, 3, 2, 2, false, true>(p, num_sms, stream);
      case 256: return launch_conv3<256, 10, 2, 3, 1, 1, false, true>(p, num_sms, stream);
      default: return cudaErrorInvalidValue;
    }
  }
  if (p.pool_out != nullptr || p.head_logits != nullptr) {   // eval-mode extras: the instantiations that carry them
    switch (block_n) {
      case 64: return resident ? launch_conv3<64, 10, 4, 3, 2, 2, true, false, true>(p, num_sms, stream)
                               : launch_conv3<64, 10, 4, 3, 2, 2, false, false, true>(p, num_sms, stream);
      case 128: return launch_conv3<128, 10, 4, 3, 2, 2, false, false, true>(p, num_sms, stream);
      case 256: return launch_conv3<256, 10, 2, 3, 1, 1, false, false, true>(p, num_sms, stream);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (block_n) {
    case 64: return resident ? launch_conv3<64, 10, 4, 3, 2, 2, true, false>(p, num_sms, stream)
                             : launch_conv3<64, 10, 4, 3, 2, 2, false, false>(p, num_sms, stream);
    case 128: return launch_conv3<128, 10, 4, 3, 2, 2, false, false>(p, num_sms, stream);
    case 256: return launch_conv3<256, 10, 2, 3, 1, 1, false, false>(p, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

template <int BLOCK_N, int S, int NSTG, int EG>
static cudaError_t launch_pix2(const PixGemmParams& p, int num_sms, cudaStream_t stream) {
  using L = Pix2Layout<BLOCK_N, S, NSTG, EG>;
  static_assert(L::kDyn <= 232448, "shared memory budget exceeded");
  auto kern = pix_gemm2_kernel<BLOCK_N, S, NSTG, EG>;
  static std::atomic<unsigned long long> attr_done{0};   // per device: function attributes belong to the context
  {
    cudaError_t ae = ensure_dynamic_smem(kern, L::kDyn, attr_done);
    if (ae != cudaSuccess) return ae;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * p.batch;
  const int units = ((m_tiles + 1) / 2) * p.n_blocks;
  if (units <= 0) return cudaSuccess;
  const int clusters = units < num_sms / 2 ? units : num_sms / 2;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(2 * clusters);
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.blockDim = dim3(L::kThreadsTotal);
  cfg.dynamicSmemBytes = L::kDyn;
  cfg.stream = stream;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return e;
  return launched();
}

cudaError_t launch_pix_gemm(const PixGemmParams& p, int block_n, int num_sms, cudaStream_t stream) {
  if (!p.pair) return cudaErrorInvalidValue;
  if (p.stem_x) return launch_stem_gemm(p, num_sms, stream);
  if (p.conv3) return launch_conv3_gemm(p, block_n, num_sms, stream);
  switch (block_n) {
    case 64: return launch_pix2<64, 5, 2, 2>(p, num_sms, stream);
    case 128: return launch_pix2<128, 4, 2, 2>(p, num_sms, stream);
    case 256: return launch_pix2<256, 3, 1, 1>(p, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

// =============================================================================================
//                                   weight-gradient GEMM
// =============================================================================================
template <int BLOCK_N, int STAGES, bool XF = false>
struct WgLayout {
  static constexpr int kDY = 2 * 16 * kAtomBytes;                 // two 64-channel blocks of 128 pixels
  static constexpr int kX = (BLOCK_N / 64) * kASlotBytes;         // BLOCK_N/64 blocks of (16+2) rows
  static constexpr int kStageSz = kDY + kX;
  static constexpr int kBar = STAGES * kStageSz;
  static constexpr int kNumBar = 2 * STAGES + 1 + (XF ? STAGES : 0);
  static constexpr int kTmemPtr = kBar + kNumBar * 8;
  static constexpr int kTotal = kTmemPtr + 16;
  static constexpr int kDyn = kTotal + 1024;
  static constexpr int kTmemCols = 3 * BLOCK_N <= 128 ? 128 : (3 * BLOCK_N <= 256 ? 256 : 512);
};

// XF: X is the raw output of the previous convolution; warps 2..5 (idle until the final accumulator drain) apply its
// BatchNorm + ReLU to the X blocks of every stage in shared memory (bnrelu_patch) between the TMA load and the MMAs.
static constexpr int kWgXfThreads = kThreads + 128;             // XF: warps 2..9 transform X; warps 2..5 drain the accumulator
template <int BLOCK_N, int STAGES, bool XF>
__global__ void __launch_bounds__(XF ? kWgXfThreads : kThreads, 1) wgrad_gemm_kernel(const __grid_constant__ WgradParams p) {
  using L = WgLayout<BLOCK_N, STAGES, XF>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* land = tmem_full + 1;                                  // XF: the stage's TMA loads have landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item: blockIdx = ((split * G + g) * n_blocks + nb) * m_blocks + mb
  int id = blockIdx.x;
  const int mb = id % p.m_blocks; id /= p.m_blocks;
  const int nb = id % p.n_blocks; id /= p.n_blocks;
  const int g = id % p.G;
  const int split = id / p.G;
  const int total_tiles = p.tiles_w * p.tiles_h * p.batch;
  const int t_begin = (int)(((long long)total_tiles * split) / p.splits);
  const int t_end = (int)(((long long)total_tiles * (split + 1)) / p.splits);
  const int per_img = p.tiles_w * p.tiles_h;
  const int m_valid = (p.Mtot - mb * 128) < 128 ? (p.Mtot - mb * 128) : 128;   // 64 or 128
  const int dy_blocks = m_valid > 64 ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], XF ? 8 : 1); mbar_init(&empty[i], 1); }
    if (XF) for (int i = 0; i < STAGES; ++i) mbar_init(&land[i], 1);
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tmapDY[p.dy_map[g]]);
    tma_prefetch_desc(&p.tmapX[p.x_map[g]]);
  }
  if (warp == 2) tmem_alloc<L::kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    int s = 0, ph = 0;
    const uint32_t x_rows_bytes = (16 + p.R - 1) * kAtomBytes;
    const uint32_t bytes = dy_blocks * 16 * kAtomBytes + (BLOCK_N / 64) * x_rows_bytes;
    for (int t = t_begin; t < t_end; ++t) {
      const int b = t / per_img, rem = t - b * per_img;
      const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
      const int w0 = tw * 8, h0 = th * 16;
      mbar_wait(&empty[s], ph ^ 1);
      if (lane == 0) {
        uint8_t* st = smem + s * L::kStageSz;
        uint64_t* bar = XF ? &land[s] : &full[s];
        mbar_arrive_expect_tx(bar, bytes);
        for (int j = 0; j < dy_blocks; ++j)
          tma_load_4d(st + j * 16 * kAtomBytes, &p.tmapDY[p.dy_map[g]], bar, p.dy_chan0 + mb * 128 + j * 64, w0,
                      h0, b);
        for (int j = 0; j < BLOCK_N / 64; ++j)
          tma_load_4d(st + L::kDY + j * kASlotBytes, &p.tmapX[p.x_map[g]], bar,
                      p.x_chan0 + nb * BLOCK_N + j * 64, w0 + p.x_dw[g], h0 + p.x_dh[g], b);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // Every lane runs this (warp-uniform) code; one elected lane issues each tcgen05 instruction.
    // One MMA covers the R vertical taps of a 64-channel block of X at once: N = 64*R, the N-blocks of the MN-major
    // B operand are the SAME 64 channels one image row (= one 1024-byte atom) further down (LBO = 1024), so the dY
    // operand is read from shared memory once per K-step instead of once per tap.
    const uint32_t idesc = p.R == 3 ? make_idesc(128, 192, 1, 1) : make_idesc(128, 64, 1, 1);
    int s = 0, ph = 0;
    uint32_t accumulate = 0;
    // With a single 64-channel block of dY (Cout == 64) both halves of the M=128 operand alias the
    // same block (LBO = 0); rows 64..127 of the accumulator are duplicates and never stored.
    const uint32_t a_lbo = dy_blocks == 2 ? 16 * kAtomBytes : 0;
    const uint64_t a_full = make_smem_desc(0, a_lbo, kAtomBytes), b_full = make_smem_desc(0, kAtomBytes, kAtomBytes);
    const uint32_t a_lbo_lo = (uint32_t)a_full, b_lbo_lo = (uint32_t)b_full;   // LBO fields (start address = 0)
    const uint32_t base = smem_u32(smem) >> 4;
    const uint32_t acc_cols = 64 * p.R;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint32_t dy_lo = base + (uint32_t)(s * (L::kStageSz >> 4));
      const uint32_t x_lo = dy_lo + (L::kDY >> 4);
#pragma unroll
      for (int j = 0; j < BLOCK_N / 64; ++j)   // eight K-steps of 16 pixels: two 1024-byte atoms (= 128 descriptor units) per step
        umma_bf16_steps_warp<false, 8, 2 * (kAtomBytes >> 4)>(tmem_base + j * acc_cols, dy_lo | a_lbo_lo,
                                                             (x_lo + j * (kASlotBytes >> 4)) | b_lbo_lo, idesc,
                                                             accumulate);
      accumulate = 1;
      umma_commit_warp<false>(&empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    umma_commit_warp<false>(tmem_full);
  } else {
    if (XF) {
      const int tt = threadIdx.x - 64;                               // 0..255
      int s = 0, ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int b = t / per_img, rem = t - b * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        mbar_wait(&land[s], ph);
        uint8_t* xs = smem + s * L::kStageSz + L::kDY;
#pragma unroll 1
        for (int j = 0; j < BLOCK_N / 64; ++j)
          bnrelu_patch<8, 256, 144, 4>(xs + j * kASlotBytes, tt, p.x_scale + nb * BLOCK_N + j * 64, p.x_shift + nb * BLOCK_N + j * 64,
                               th * 16 + p.x_dh[g], tw * 8 + p.x_dw[g], p.H, p.W, true);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
    if (warp < 6) {
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (t_end > t_begin) {
      const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
      for (int j = 0; j < BLOCK_N / 64; ++j) {
        for (int r = 0; r < p.R; ++r) {
          float* dst_row = p.dw + ((size_t)((g * p.R + r) * p.Mtot + mb * 128 + row)) * p.Ntot + nb * BLOCK_N + j * 64;
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + (j * p.R + r) * 64 + c0, v);
            tmem_ld_wait();
            if (row < m_valid) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + c0 + 4 * q),
                             "f"(__uint_as_float(v[4 * q])), "f"(__uint_as_float(v[4 * q + 1])),
                             "f"(__uint_as_float(v[4 * q + 2])), "f"(__uint_as_float(v[4 * q + 3]))
                             : "memory");
              }
            }
          }
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<L::kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N, int STAGES, bool XF>
static cudaError_t launch_wg(const WgradParams& p, cudaStream_t stream) {
  using L = WgLayout<BLOCK_N, STAGES, XF>;
  static_assert(L::kDyn <= 232448, "shared memory budget exceeded");
  auto kern = wgrad_gemm_kernel<BLOCK_N, STAGES, XF>;
  static std::atomic<unsigned long long> attr_done{0};   // per device: function attributes belong to the context
  {
    cudaError_t ae = ensure_dynamic_smem(kern, L::kDyn, attr_done);
    if (ae != cudaSuccess) return ae;
  }
  const int grid = p.m_blocks * p.n_blocks * p.G * p.splits;
  if (grid <= 0) return cudaSuccess;
  kern<<<grid, XF ? kWgXfThreads : kThreads, L::kDyn, stream>>>(p);
  return launched();
}

// =============================================================================================
//                     weight gradient of a 3x3 convolution with Cout == 64: nine taps per CTA
// =============================================================================================
// wgrad_gemm_kernel puts Cout on the M = 128 rows: with Cout = 64 half of every MMA is wasted, and its three horizontal
// taps live in three CTAs that each re-read dY and an 8 x 18 patch of X (the kernel runs at the L2 -> SM bandwidth cap).
// Here the roles are swapped: D[(tap, cin)][cout] = sum_pixels X[pixel + tap, cin] * dY[pixel, cout].  ONE (8+2) x (16+2)
// patch of X serves all nine taps (tap (r, g) = row offset r*10 + g, row-group stride 1280 B — cf. conv3_gemm_kernel); an
// M = 128 MMA covers two taps (the second 64-row block is the same patch `LBO` bytes further: +128 B = next horizontal tap,
// +1280 B = next vertical tap), so nine taps are five MMAs per 16-pixel K-step (the last with a duplicated block), N = 64.
// Per 128-pixel tile: 39 KB of loads for 9 taps x 64 x 64 outputs, instead of 3 x 34 KB.
template <int STAGES, bool XF = false>
struct Wg9Layout {
  static constexpr int kDY = 16 * kAtomBytes;                     // 128 pixels x 64 output channels
  static constexpr int kXPatch = 18 * 10 * 128;
  static constexpr int kX = (kXPatch + 1023) & ~1023;
  static constexpr int kStageSz = kDY + kX;
  static constexpr int kBar = STAGES * kStageSz;
  static constexpr int kNumBar = 2 * STAGES + 1 + (XF ? STAGES : 0);
  static constexpr int kTmemPtr = kBar + kNumBar * 8;
  static constexpr int kTotal = kTmemPtr + 16;
  static constexpr int kDyn = kTotal + 1024;
  static constexpr int kTmemCols = 512;                           // five 64-column accumulators
};

template <int STAGES, bool XF>
__global__ void __launch_bounds__(XF ? kWgXfThreads : kThreads, 1) wgrad9_gemm_kernel(const __grid_constant__ WgradParams p) {
  using L = Wg9Layout<STAGES, XF>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* land = tmem_full + 1;                                  // XF: the stage's TMA loads have landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item: blockIdx = split * n_blocks + nb   (nb = 64-channel block of Cin)
  const int nb = blockIdx.x % p.n_blocks;
  const int split = blockIdx.x / p.n_blocks;
  const int total_tiles = p.tiles_w * p.tiles_h * p.batch;
  const int t_begin = (int)(((long long)total_tiles * split) / p.splits);
  const int t_end = (int)(((long long)total_tiles * (split + 1)) / p.splits);
  const int per_img = p.tiles_w * p.tiles_h;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], XF ? 8 : 1); mbar_init(&empty[i], 1); }
    if (XF) for (int i = 0; i < STAGES; ++i) mbar_init(&land[i], 1);
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tmapDY[0]);
    tma_prefetch_desc(&p.tmapX9);
  }
  if (warp == 2) tmem_alloc<L::kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    int s = 0, ph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int b = t / per_img, rem = t - b * per_img;
      const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
      const int w0 = tw * 8, h0 = th * 16;
      mbar_wait(&empty[s], ph ^ 1);
      if (lane == 0) {
        uint8_t* st = smem + s * L::kStageSz;
        uint64_t* bar = XF ? &land[s] : &full[s];
        mbar_arrive_expect_tx(bar, L::kDY + L::kXPatch);
        tma_load_4d(st, &p.tmapDY[0], bar, p.dy_chan0, w0, h0, b);
        tma_load_4d(st + L::kDY, &p.tmapX9, bar, p.x_chan0 + nb * 64, w0 - 1, h0 - 1, b);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // A = X patch (MN-major: rows = pixels, 64 input channels per 128-byte row), M = 128 = two taps; B = dY (MN-major).
    constexpr uint32_t idesc = make_idesc(128, 64, 1, 1);
    constexpr uint32_t a_hi = ((uint32_t)(10 * 128) >> 4) | (1u << 14) | (2u << 29);   // SBO = one patch row group
    constexpr uint32_t b_hi = kDescHiSw128;
    const uint32_t base = smem_u32(smem) >> 4;
    // accumulator q: first tap (r, g) and the distance to the second one
    //   q0..q2: (r, 0) + (r, 1)   LBO = 128 B      q3: (0, 2) + (1, 2)   LBO = 1280 B      q4: (2, 2), duplicated (LBO = 0)
    int s = 0, ph = 0;
    uint32_t accumulate = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint32_t dy_lo = (base + (uint32_t)(s * (L::kStageSz >> 4))) | ((1024u >> 4) << 16);
      const uint32_t x_lo = base + (uint32_t)((s * L::kStageSz + L::kDY) >> 4);
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const uint32_t row0 = q < 3 ? (uint32_t)(q * 10) : (q == 3 ? 2u : 22u);
        const uint32_t lbo = q < 3 ? 128u : (q == 3 ? 1280u : 0u);
        // eight K-steps of 16 pixels: two image rows of the patch (2 x 1280 B) / of the dY tile (2 x 1024 B) per step
        umma_bf16_steps8_warp_ab<2 * 1280 / 16, 2 * 1024 / 16>(tmem_base + q * 64, (x_lo + row0 * 8u) | ((lbo >> 4) << 16), a_hi,
                                                               dy_lo, b_hi, idesc, accumulate);
      }
      accumulate = 1;
      umma_commit_warp<false>(&empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    umma_commit_warp<false>(tmem_full);
  } else {
    if (XF) {
      const int tt = threadIdx.x - 64;                               // 0..255
      int s = 0, ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int b = t / per_img, rem = t - b * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        mbar_wait(&land[s], ph);
        bnrelu_patch<10, 256, 180, 5>(smem + s * L::kStageSz + L::kDY, tt, p.x_scale + nb * 64, p.x_shift + nb * 64, th * 16 - 1,
                              tw * 8 - 1, p.H, p.W, true);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
    if (warp < 6) {
    const int sub = warp & 3;
    const int row = sub * 32 + lane;           // accumulator row = (which tap of the pair) * 64 + input channel
    const int half = row >> 6, cin = row & 63;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (t_end > t_begin) {
      const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
#pragma unroll 1
      for (int q = 0; q < 5; ++q) {
        // packed tap index = g * 3 + r
        int r, g;
        if (q < 3) { r = q; g = half; }
        else if (q == 3) { r = half; g = 2; }
        else { r = 2; g = 2; }
        const bool live = !(q == 4 && half == 1);
        float* dst = p.dw + ((size_t)(g * 3 + r) * p.Mtot) * p.Ntot + nb * 64 + cin;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + q * 64 + c0, v);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (size_t)(c0 + i) * p.Ntot), "f"(__uint_as_float(v[i])) : "memory");
          }
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<L::kTmemCols>(tmem_base);
  }
}

template <bool XF>
static cudaError_t launch_wgrad9(const WgradParams& p, cudaStream_t stream) {
  constexpr int STAGES = 5;
  using L = Wg9Layout<STAGES, XF>;
  static_assert(L::kDyn <= 232448, "shared memory budget exceeded");
  auto kern = wgrad9_gemm_kernel<STAGES, XF>;
  static std::atomic<unsigned long long> attr_done{0};
  {
    cudaError_t ae = ensure_dynamic_smem(kern, L::kDyn, attr_done);
    if (ae != cudaSuccess) return ae;
  }
  if (p.Mtot != 64 || p.G != 3 || p.R != 3) return cudaErrorInvalidValue;
  const int grid = p.n_blocks * p.splits;
  if (grid <= 0) return cudaSuccess;
  kern<<<grid, XF ? kWgXfThreads : kThreads, L::kDyn, stream>>>(p);
  return launched();
}

cudaError_t launch_wgrad_gemm(const WgradParams& p, int block_n, cudaStream_t stream) {
  const bool xf = p.x_scale && p.x_shift;
  if (xf && (p.R != 3 || p.G != 3)) return cudaErrorInvalidValue;    // the X transform is built for the 3x3 convolutions
  if (p.nine) return xf ? launch_wgrad9<true>(p, stream) : launch_wgrad9<false>(p, stream);
  switch (block_n) {
    case 64: return xf ? launch_wg<64, 4, true>(p, stream) : launch_wg<64, 4, false>(p, stream);
    case 128: return xf ? launch_wg<128, 3, true>(p, stream) : launch_wg<128, 3, false>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}

// =============================================================================================
//                                   tensor-map encoding
// =============================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

int make_tmap_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                 const uint32_t box[4]) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t e[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), d, s, b, e,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

int make_stem_tmap(CUtensorMap* out, const float* x, int B, int C, int H, int W) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t d[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t s[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
  cuuint32_t b[4] = {(cuuint32_t)kStemPatchW, (cuuint32_t)kStemPatchH, (cuuint32_t)C, 1};
  cuuint32_t e[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t d[2] = {inner, rows};
  cuuint64_t s[1] = {row_stride_bytes};
  cuuint32_t b[2] = {box_inner, box_rows};
  cuuint32_t e[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), d, s, b, e,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

}  // namespace cs
