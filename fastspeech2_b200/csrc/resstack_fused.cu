// One persistent tcgen05 kernel per HiFi-GAN upsample stage for the multi-receptive-field ResBlock group
//
//   y = (1/n_kernels) * sum_j ResBlock_j(x),   ResBlock_j: x <- conv_{k_j,1}(lrelu(conv_{k_j,d}(lrelu(x)))) + x  for d in dilations
//
// (hifigan/models.py:154-160 and ResBlock.forward :96-103; contract: fs2_resstack in include/fs2b200.h).  The per-layer design
// issued 18 conv launches per stage and moved every intermediate through HBM (~48 full-tensor passes); here a CTA owns one time
// tile of one utterance and keeps the whole chain on chip:
//
//   * slab = MT*128 rows starting H = max_j 6(k_j-1)... (sum of the receptive radii) rows before the tile: every conv is evaluated on
//     the full slab, garbage from the slab edges grows inward by one conv radius per layer and by construction stays inside the
//     H-row halo (halo recompute); rows outside the utterance [0, N) are forced to zero after every layer (Conv1d zero padding).
//   * activations live in shared memory ALREADY in tensor-core operand form: per 16-channel K-block an fp16 "hi" plane and an
//     E4M3 plane [lo * 2^12 | hi * 2] (the two-MMA operand split of conv_tc_kernel.cuh, FS2_TC_VARIANT_F8), UMMA no-swizzle
//     K-major [16-byte K-chunk][row][16 B], so a conv tap is the same slab with the descriptor start advanced by tap*dilation
//     rows.  Two slabs: XA = lrelu(x) (conv1's operand), XT = lrelu(conv1 output) (conv2's operand).
//   * the residual stream x stays in TENSOR MEMORY in fp32 (MT*C columns) next to the MT accumulators (MT*C columns):
//     epilogue = tcgen05.ld acc (+ tcgen05.ld x) -> bias / residual / lrelu -> operand split -> st.shared into the other slab
//     (+ tcgen05.st x).  Thread == slab row, so consecutive lanes write consecutive 16-byte rows: conflict-free.
//   * weights stream through a cp.async.bulk ring exactly as in conv_tc_kernel.cuh (same tile images: the packer's f8 format).
//   * HBM: the tile of x is read (once from HBM, again from L2 for the other kernel sizes), y is written / accumulated in L2.
//
// Roles: warp 0 weight producer, warp 1 MMA issuer, warps 2-9 "row" warps (TMEM lane quarter = warp % 4, column half = (warp-2)/4)
// that load the tile, run every epilogue and store the result.
#include "conv_tc_kernel.cuh"

namespace fs2 {

constexpr int RS_MAXK = FS2_MAX_DIL + 4;   // kernel sizes per stage
constexpr int RS_THREADS = 320;
constexpr int RS_GUARD = 1024;             // zeroed bytes in front of the first slab (taps reach up to 32 rows before row 0)
constexpr int RS_SB_MAX = 16;

struct RsConv { const unsigned char* w; const float* b; int taps, dil; };
struct RsP {
  const float* x; float* y;
  int B, N;
  int n_kernels, n_dil;
  RsConv conv[RS_MAXK][FS2_MAX_DIL][2];
  int H, TILE, tiles_per_b, n_items;
  float alpha;
  int SB;
};

__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 channels of one slab row -> operand planes of K-block kb (fp16 hi: two 16-byte chunks; E4M3: [lo*2^12 | hi*2]).
// `a` already carries the activation and the out-of-utterance zeroing.
__device__ __forceinline__ void rs_store16(unsigned char* kblk, uint32_t chunk_bytes, int row, const float (&a)[16]) {
  uint32_t hw[8], l8[4], h8[4];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const float a0 = a[2 * j], a1 = a[2 * j + 1];
    hw[j] = cvt_f16x2_sat(a0, a1);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
    const uint32_t l = cvt_e4m3x2_sat((a0 - hf.x) * TC_F8_LO_SCALE, (a1 - hf.y) * TC_F8_LO_SCALE);
    const uint32_t h = cvt_e4m3x2_sat(hf.x * TC_F8_HI_SCALE, hf.y * TC_F8_HI_SCALE);
    if (j & 1) { l8[j >> 1] |= l << 16; h8[j >> 1] |= h << 16; }
    else { l8[j >> 1] = l; h8[j >> 1] = h; }
  }
  unsigned char* p0 = kblk + (size_t)row * 16;
  *reinterpret_cast<uint4*>(p0) = make_uint4(hw[0], hw[1], hw[2], hw[3]);                       // plane 0, chunk 0: channels 0-7
  *reinterpret_cast<uint4*>(p0 + chunk_bytes) = make_uint4(hw[4], hw[5], hw[6], hw[7]);         // plane 0, chunk 1: channels 8-15
  *reinterpret_cast<uint4*>(p0 + 2 * chunk_bytes) = make_uint4(l8[0], l8[1], l8[2], l8[3]);     // plane 1, chunk 0: E4M3 lo
  *reinterpret_cast<uint4*>(p0 + 3 * chunk_bytes) = make_uint4(h8[0], h8[1], h8[2], h8[3]);     // plane 1, chunk 1: E4M3 hi
}

__device__ __forceinline__ float rs_lrelu(float v) { return fmaxf(v, 0.1f * v); }   // LRELU_SLOPE = 0.1 (hifigan/models.py:7)

template <int C, int MT>
__global__ void __launch_bounds__(RS_THREADS, 1) resstack_kernel(const RsP p) {
  constexpr int KB = C / 16, R = MT * 128, HC = C / 2, NG = HC / 16;   // NG: 16-channel groups per row warp
  constexpr uint32_t CHUNK = (uint32_t)R * 16, PLANE = 2 * CHUNK, KBLK = 2 * PLANE, SLAB = KB * KBLK;
  constexpr uint32_t WSTAGE = 64u * C;
  constexpr uint32_t TMEM_COLS = (2 * MT * C) <= 256 ? 256 : 512;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* xa = smem_raw + RS_GUARD;
  unsigned char* xt = xa + SLAB;
  unsigned char* ring = xt + SLAB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.SB * WSTAGE);
  uint64_t* fullB = bars;                       // [RS_SB_MAX]
  uint64_t* emptyB = fullB + RS_SB_MAX;         // [RS_SB_MAX]
  uint64_t* accFull = emptyB + RS_SB_MAX;       // [MT] MMAs of one conv for tile m have retired
  uint64_t* rowsReady = accFull + 4;            // [MT] all 8 row warps have produced tile m of the next conv's operand slab (and drained its accumulator)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rowsReady + 4);

  for (int i = tid; i < RS_GUARD / 16; i += RS_THREADS) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < RS_SB_MAX; i++) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < 4; i++) { mbar_init(&accFull[i], 1); mbar_init(&rowsReady[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t acc_col = 0, x_col = (uint32_t)(MT * C);     // accumulators | residual stream

  const int rounds_per_item = p.n_kernels;
  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      Ring rb;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x)
        for (int j = 0; j < p.n_kernels; j++)
          for (int d = 0; d < p.n_dil; d++)
            for (int c2 = 0; c2 < 2; c2++) {
              const RsConv cv = p.conv[j][d][c2];
              const unsigned char* src = cv.w + TC_HDR;
              const int stages = KB * cv.taps;
              for (int s = 0; s < stages; s++) {
                mbar_wait(&emptyB[rb.idx], rb.phase ^ 1);
                mbar_expect_tx(&fullB[rb.idx], WSTAGE);
                bulk_g2s(ring + (size_t)rb.idx * WSTAGE, src, WSTAGE, &fullB[rb.idx]);
                src += WSTAGE;
                rb.advance((uint32_t)p.SB);
              }
            }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t idesc = umma_idesc_f16(C);
    const uint64_t a_const = umma_desc(0, CHUNK, 128), b_const = umma_desc(0, (uint32_t)C * 16, 128);
    const uint32_t xa16 = smem_u32(xa) >> 4, xt16 = smem_u32(xt) >> 4;
    Ring rb;
    uint32_t ev = 0;                                        // events every tile's rowsReady barrier has completed so far (as waited here)
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      for (int j = 0; j < rounds_per_item; j++) {
        for (int d = 0; d < p.n_dil; d++) {
          for (int c2 = 0; c2 < 2; c2++) {
            const RsConv cv = p.conv[j][d][c2];
            const int pad = (cv.taps - 1) * cv.dil / 2;
            // operand slab of this conv complete for every tile (load phase / previous epilogue), accumulators drained
#pragma unroll
            for (int m = 0; m < MT; m++) mbar_wait(&rowsReady[m], ev & 1);
            ev++;
            tc_fence_after();
            const uint32_t slab16 = c2 == 0 ? xa16 : xt16;
            for (int kb = 0; kb < KB; kb++) {
              const uint64_t a_hi = a_const | (uint64_t)((slab16 + kb * (KBLK >> 4)) & 0x3fff);
              int row_off = -pad;
              for (int tap = 0; tap < cv.taps; tap++, row_off += cv.dil, rb.advance((uint32_t)p.SB)) {
                mbar_wait(&fullB[rb.idx], rb.phase);
                tc_fence_after();
                if (leader) {
                  const uint64_t b_hi = b_const | (uint64_t)(smem_u32(ring + (size_t)rb.idx * WSTAGE) >> 4);
                  const uint64_t b_x8 = b_hi + ((2u * C * 16u) >> 4);
                  const uint64_t ah0 = a_hi + (uint64_t)(int64_t)row_off;        // start-address field += rows (16 B each); never carries out of the field
                  const uint64_t ax0 = ah0 + (PLANE >> 4);
                  const uint32_t first = (kb | tap) ? 1u : 0u;
#pragma unroll
                  for (int m = 0; m < MT; m++) tc_mma_f16(tmem + acc_col + m * C, ah0 + m * 128, b_hi, idesc, first);
#pragma unroll
                  for (int m = 0; m < MT; m++) tc_mma_f8(tmem + acc_col + m * C, ax0 + m * 128, b_x8, idesc, 1u);
                  tc_commit(&emptyB[rb.idx]);
                }
                __syncwarp();
              }
            }
            if (leader) {
#pragma unroll
              for (int m = 0; m < MT; m++) tc_commit(&accFull[m]);
            }
            __syncwarp();
          }
        }
        // The round's final epilogue does not signal: a row warp's next arrival on rowsReady[m] is its load of the NEXT round, which it
        // reaches only after finishing that epilogue (program order), so every barrier phase is waited here exactly once before the
        // next one can complete (no parity aliasing).
      }
    }
  } else {
    // ===================== row warps: load, epilogues, store =====================
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int col0 = h * HC;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int b = item / p.tiles_per_b;
      const int t0 = (item - b * p.tiles_per_b) * p.TILE;
      const float* xb = p.x + (size_t)b * p.N * C;
      float* yb = p.y + (size_t)b * p.N * C;
      for (int j = 0; j < p.n_kernels; j++) {
        // ---- load the slab of x: residual stream -> TMEM, lrelu(x) -> XA operand planes
#pragma unroll 1
        for (int m = 0; m < MT; m++) {
          const int row = m * 128 + q * 32 + lane;
          const int g = t0 - p.H + row;
          const bool in = g >= 0 && g < p.N;
#pragma unroll
          for (int gi = 0; gi < NG; gi++) {
            float v[16];
            if (in) {
              const float4* src = reinterpret_cast<const float4*>(xb + (size_t)g * C + col0 + gi * 16);
#pragma unroll
              for (int k4 = 0; k4 < 4; k4++) {
                const float4 u = __ldg(src + k4);
                v[4 * k4] = u.x; v[4 * k4 + 1] = u.y; v[4 * k4 + 2] = u.z; v[4 * k4 + 3] = u.w;
              }
            } else {
#pragma unroll
              for (int k = 0; k < 16; k++) v[k] = 0.f;
            }
            uint32_t raw[16];
#pragma unroll
            for (int k = 0; k < 16; k++) raw[k] = __float_as_uint(v[k]);
            tc_st16(tmem + lane_base + x_col + m * C + col0 + gi * 16, raw);
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = rs_lrelu(v[k]);
            rs_store16(xa + (size_t)((col0 >> 4) + gi) * KBLK, CHUNK, row, v);
          }
          tc_wait_st();
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&rowsReady[m]);
        }
        for (int d = 0; d < p.n_dil; d++) {
          const bool last = d == p.n_dil - 1;
          for (int c2 = 0; c2 < 2; c2++) {
            const RsConv cv = p.conv[j][d][c2];
            const float inv_s = __ldg(reinterpret_cast<const float*>(cv.w));
#pragma unroll 1
            for (int m = 0; m < MT; m++) {
              const int row = m * 128 + q * 32 + lane;
              const int g = t0 - p.H + row;
              const bool in = g >= 0 && g < p.N;
              mbar_wait(&accFull[m], acc_phase);
              tc_fence_after();
#pragma unroll
              for (int gi = 0; gi < NG; gi++) {
                const int cc = col0 + gi * 16;
                uint32_t av[16], xv[16];
                tc_ld16_nowait(tmem + lane_base + acc_col + m * C + cc, av);
                if (c2 == 1) tc_ld16_nowait(tmem + lane_base + x_col + m * C + cc, xv);
                tc_wait_ld();
                float v[16];
                const float4* bp = reinterpret_cast<const float4*>(cv.b + cc);
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++) {
                  const float4 bb = __ldg(bp + k4);
                  v[4 * k4] = fmaf(__uint_as_float(av[4 * k4]), inv_s, bb.x);
                  v[4 * k4 + 1] = fmaf(__uint_as_float(av[4 * k4 + 1]), inv_s, bb.y);
                  v[4 * k4 + 2] = fmaf(__uint_as_float(av[4 * k4 + 2]), inv_s, bb.z);
                  v[4 * k4 + 3] = fmaf(__uint_as_float(av[4 * k4 + 3]), inv_s, bb.w);
                }
                if (c2 == 0) {
                  // conv1: lrelu -> conv2's operand slab
#pragma unroll
                  for (int k = 0; k < 16; k++) v[k] = in ? rs_lrelu(v[k]) : 0.f;
                  rs_store16(xt + (size_t)(cc >> 4) * KBLK, CHUNK, row, v);
                } else {
#pragma unroll
                  for (int k = 0; k < 16; k++) v[k] += __uint_as_float(xv[k]);          // + residual
                  if (!last) {
                    uint32_t raw[16];
#pragma unroll
                    for (int k = 0; k < 16; k++) raw[k] = __float_as_uint(v[k]);
                    tc_st16(tmem + lane_base + x_col + m * C + cc, raw);
#pragma unroll
                    for (int k = 0; k < 16; k++) v[k] = in ? rs_lrelu(v[k]) : 0.f;
                    rs_store16(xa + (size_t)(cc >> 4) * KBLK, CHUNK, row, v);
                  } else if (in && row >= p.H && row < p.H + p.TILE) {
                    // result of this kernel size: y = (j ? y : 0) + alpha * x   (mean over kernel sizes, models.py:154-160)
                    float4* dst = reinterpret_cast<float4*>(yb + (size_t)g * C + cc);
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++) {
                      float4 o = make_float4(v[4 * k4] * p.alpha, v[4 * k4 + 1] * p.alpha, v[4 * k4 + 2] * p.alpha, v[4 * k4 + 3] * p.alpha);
                      if (j > 0) {
                        const float4 prev = dst[k4];
                        o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
                      }
                      dst[k4] = o;
                    }
                  }
                }
              }
              if (c2 == 1 && last) {
                tc_fence_before();                     // accumulator / residual reads done; the next signal is the next round's load
              } else {
                if (c2 == 1) tc_wait_st();
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&rowsReady[m]);
              }
            }
            acc_phase ^= 1;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------ host side
static size_t rs_smem_bytes(int C, int MT, int SB) {
  const size_t slab = (size_t)(C / 16) * 4 * (MT * 128) * 16;
  return RS_GUARD + 2 * slab + (size_t)SB * 64 * C + (2 * RS_SB_MAX + 8) * 8 + 16;
}

// Launch plan (pure host logic): out[8] = {MT, H (halo rows per side), TILE (output rows per work item), work items, grid, weight ring
// stages, dynamic shared memory bytes, TMEM columns}
int resstack_plan(const fs2_resstack_args* a, int num_sms, int* out) {
  if (!a || a->B <= 0 || a->N <= 0 || num_sms <= 0) return FS2_ERR_ARG;
  if (a->C != 32 && a->C != 64) return FS2_ERR_UNSUPPORTED;
  if (a->n_kernels <= 0 || a->n_kernels > RS_MAXK || a->n_dil <= 0 || a->n_dil > FS2_MAX_DIL) return FS2_ERR_ARG;
  int H = 0;
  for (int j = 0; j < a->n_kernels; j++) {
    const int k = a->k[j];
    if (k <= 0 || !(k & 1)) return FS2_ERR_UNSUPPORTED;
    int hj = 0;
    for (int d = 0; d < a->n_dil; d++) {
      const int dil = a->dil[j][d];
      if (dil <= 0 || (k - 1) * dil / 2 > 32) return FS2_ERR_UNSUPPORTED;      // taps reach at most 32 rows outside a tile (guard / neighbour tile)
      hj += (k - 1) * dil / 2 + (k - 1) / 2;
    }
    H = hj > H ? hj : H;
  }
  const int MT = a->C == 32 ? 4 : 3;            // slab rows = MT*128: bounded by shared memory (two slabs of 4*C bytes per row)
  const int TILE = MT * 128 - 2 * H;
  if (TILE < 64) return FS2_ERR_UNSUPPORTED;
  const long long tiles_per_b = (a->N + TILE - 1) / TILE, items = tiles_per_b * a->B;
  if (items > 0x7fffffffLL) return FS2_ERR_UNSUPPORTED;
  int SB = RS_SB_MAX;
  while (SB > 4 && rs_smem_bytes(a->C, MT, SB) > 226 * 1024) SB--;
  if (rs_smem_bytes(a->C, MT, SB) > 226 * 1024) return FS2_ERR_UNSUPPORTED;
  out[0] = MT; out[1] = H; out[2] = TILE; out[3] = (int)items; out[4] = items < num_sms ? (int)items : num_sms; out[5] = SB;
  out[6] = (int)rs_smem_bytes(a->C, MT, SB); out[7] = 2 * MT * a->C <= 256 ? 256 : 512;
  return FS2_OK;
}

int resstack(const fs2_resstack_args* a, cudaStream_t s) {
  if (!a || !a->x || !a->y) return FS2_ERR_ARG;
  if (!aligned16(a->x) || !aligned16(a->y)) return FS2_ERR_ARG;
  int derr = FS2_OK;
  DevState* dv = dev_state(&derr);
  if (!dv) return derr;
  int plan[8];
  FS2_TRY(resstack_plan(a, dv->num_sms.load(std::memory_order_relaxed), plan));
  if (!dv->fused_ready.load(std::memory_order_acquire)) {
    DevOnce once;
    if (!dv->fused_ready.load(std::memory_order_relaxed)) {
      const int mx = 227 * 1024;
      cudaError_t e = cudaFuncSetAttribute(resstack_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(resstack_kernel<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
      if (e != cudaSuccess) return FS2_ERR_CUDA - (int)e;
      dv->fused_ready.store(true, std::memory_order_release);
    }
  }
  RsP p{};
  p.x = a->x; p.y = a->y; p.B = a->B; p.N = a->N;
  p.n_kernels = a->n_kernels; p.n_dil = a->n_dil;
  double flops = 0;
  for (int j = 0; j < a->n_kernels; j++)
    for (int d = 0; d < a->n_dil; d++) {
      if (!a->w1_tc[j][d] || !a->w2_tc[j][d] || !a->b1[j][d] || !a->b2[j][d]) return FS2_ERR_ARG;
      if (!aligned16(a->w1_tc[j][d]) || !aligned16(a->w2_tc[j][d]) || !aligned16(a->b1[j][d]) || !aligned16(a->b2[j][d])) return FS2_ERR_ARG;
      p.conv[j][d][0] = RsConv{reinterpret_cast<const unsigned char*>(a->w1_tc[j][d]), a->b1[j][d], a->k[j], a->dil[j][d]};
      p.conv[j][d][1] = RsConv{reinterpret_cast<const unsigned char*>(a->w2_tc[j][d]), a->b2[j][d], a->k[j], 1};
      flops += 2.0 * 2.0 * a->B * (double)a->N * a->C * a->C * a->k[j];
    }
  p.H = plan[1]; p.TILE = plan[2]; p.tiles_per_b = (a->N + p.TILE - 1) / p.TILE; p.n_items = plan[3];
  p.alpha = 1.f / (float)a->n_kernels; p.SB = plan[5];
  prof_before(s);
  if (a->C == 32) resstack_kernel<32, 4><<<plan[4], RS_THREADS, plan[6], s>>>(p);
  else resstack_kernel<64, 3><<<plan[4], RS_THREADS, plan[6], s>>>(p);
  prof_after(s, 0, flops);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

}  // namespace fs2
