// tcgen05 (5th-gen tensor core) implicit-GEMM Conv1d over channels-last activations, error-compensated 3xTF32.
//
//   y[b,t,n] = epilogue( sum_{tap} sum_c act(x[b, t + tap*dil - pad, c]) * w[tap][c][n] )        (contract: fs2_conv1d)
//
// Why 3xTF32: single-pass TF32 misses the parity bars (mel 1.2e-3 vs 1e-3, waveform 5.3e-4 vs 1e-4).
// Each fp32 operand is split x = hi + lo with hi = x & 0xffffe000 (exactly a TF32 value) and lo = x - hi (exact in fp32);
// D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi with fp32 accumulation in TMEM leaves a relative error of ~2^-21 per product.
//
// Data movement per CTA (one utterance, MT consecutive 128-row time tiles, one block of N <= 256 output channels):
//   * A (activations): 4 transform warps read the [MT*128 + (taps-1)*dil] x 16-channel slab of the current K-block ONCE from
//     global (float4, coalesced), apply the input activation, split hi/lo and store both in the UMMA no-swizzle K-major
//     layout  [16-byte K-chunk][row][4 floats].  In that layout a core matrix (8 rows x 16 B) starting at ANY row is 128
//     contiguous bytes, so every conv tap is just a descriptor whose start address is advanced by tap*dil rows: the slab is
//     loaded and split once per K-block, not once per tap.
//   * B (weights): pre-split and pre-tiled on the host into the exact smem image of one (tap, K-block) stage
//     ([hi|lo][K-chunk][n][4 floats]); one cp.async.bulk (TMA bulk engine) per stage, mbarrier complete_tx.
//   * D: MT accumulators of 128 lanes x N fp32 columns in TMEM; each B stage feeds MT*2*3 MMAs (M=128, K=8).
//   * Epilogue: the transform warps turn into epilogue warps: tcgen05.ld (thread == output row) -> bias / activation /
//     residual / alpha / accumulate / pad-row mask -> 16-byte global stores.
// Warp roles: warp 0 = weight-stage producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = transform then epilogue.
#include "common.cuh"

namespace fs2 {

constexpr int TC_KB = 16;          // input channels per K-block (two K=8 TF32 MMAs)
constexpr int TC_CHUNKS = TC_KB / 4;
constexpr int TC_SA = 2;           // activation slab stages
constexpr int TC_SB_MAX = 4;       // weight stages (runtime: p.SB <= TC_SB_MAX)
constexpr int TC_THREADS = 192;
constexpr int TC_LD = 8;           // 16-byte global loads in flight per transform thread

struct TcP {
  const float* x; long long xbs, xrs;
  int B, T, Cin;
  const float* wt;                 // tiled weights, see pack_conv_tc()
  const float* bias;
  int N;                           // total output channels
  int NB;                          // output channels per CTA (MMA N), N % NB == 0, NB % 16 == 0, NB <= 256
  int taps, dil, pad;
  int in_act; float in_slope;
  int out_act; float out_slope;
  const float* res; long long rbs, rrs;
  float alpha; int accumulate;
  const int* row_lens;
  float* y; long long ybs, yrs;
  int MT;                          // 128-row tiles per CTA
  int SB;                          // weight stages in flight
  int R;                           // slab rows held in smem (>= MT*128 + (taps-1)*dil, R % 8 == 2)
  int tiles_per_batch;
  int acc_stride;                  // TMEM columns between accumulators
  int tmem_cols;                   // power of two >= MT*acc_stride
  unsigned variant;                // debug: bit0 swaps LBO/SBO
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, no-swizzle K-major: core matrix = 8 rows x 16 B stored contiguously (128 B);
// LBO = byte distance between the two 16-byte K-chunks of one K=8 (TF32) MMA, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // layout_type = SWIZZLE_NONE (0), base_offset = 0
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc_kernel(const TcP p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.R, NB = p.NB;
  const uint32_t a_plane = (uint32_t)TC_CHUNKS * R * 16;          // bytes of one hi (or lo) slab
  const uint32_t b_plane = (uint32_t)TC_CHUNKS * NB * 16;         // bytes of one hi (or lo) weight tile
  unsigned char* a_base = smem_raw;                                // [SA][hi|lo][chunk][R][16 B]
  unsigned char* b_base = a_base + (size_t)TC_SA * 2 * a_plane;    // [SB][hi|lo][chunk][NB][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)p.SB * 2 * b_plane);
  uint64_t* fullA = bars;                // [SA]
  uint64_t* emptyA = bars + TC_SA;       // [SA]
  uint64_t* fullB = emptyA + TC_SA;      // [SB_MAX]
  uint64_t* emptyB = fullB + TC_SB_MAX;  // [SB_MAX]
  uint64_t* accFull = emptyB + TC_SB_MAX;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accFull + 1);

  const int b = blockIdx.x / p.tiles_per_batch;
  const int t0 = (blockIdx.x % p.tiles_per_batch) * p.MT * 128;
  const int nblk = blockIdx.y;
  const int KBLOCKS = p.Cin / TC_KB;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < TC_SA; i++) { mbar_init(&fullA[i], 128); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < TC_SB_MAX; i++) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    mbar_init(accFull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== weight-stage producer (TMA bulk copies) =====================
    if (lane == 0) {
      const uint32_t stage_bytes = 2 * b_plane;
      const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.wt) + (size_t)nblk * p.taps * KBLOCKS * stage_bytes;
      int it = 0;
      for (int kb = 0; kb < KBLOCKS; kb++) {
        for (int tap = 0; tap < p.taps; tap++, it++) {
          const int s = it % p.SB;
          const uint32_t ph = (it / p.SB) & 1;
          mbar_wait(&emptyB[s], ph ^ 1);
          mbar_expect_tx(&fullB[s], stage_bytes);
          bulk_g2s(b_base + (size_t)s * stage_bytes, wsrc + ((size_t)tap * KBLOCKS + kb) * stage_bytes, stage_bytes, &fullB[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(NB);
      // descriptors differ only in the 14-bit start-address field (16-byte units): build the constant part once and add offsets
      const uint64_t a_const = umma_desc(0, (uint32_t)R * 16, 128), b_const = umma_desc(0, (uint32_t)NB * 16, 128);
      const uint32_t a_kk = 2u * (uint32_t)R, b_kk = 2u * (uint32_t)NB;     // two 16-byte K-chunks per K=8 MMA, in 16-byte units
      int it = 0;
      for (int kb = 0; kb < KBLOCKS; kb++) {
        const int sa = kb % TC_SA;
        mbar_wait(&fullA[sa], (kb / TC_SA) & 1);
        tc_fence_after();
        const uint64_t a_hi = a_const | (uint64_t)(smem_u32(a_base + (size_t)sa * 2 * a_plane) >> 4);
        const uint64_t a_lo = a_hi + (a_plane >> 4);
        for (int tap = 0; tap < p.taps; tap++, it++) {
          const int sb = it % p.SB;
          mbar_wait(&fullB[sb], (it / p.SB) & 1);
          tc_fence_after();
          const uint64_t b_hi = b_const | (uint64_t)(smem_u32(b_base + (size_t)sb * 2 * b_plane) >> 4);
          const uint64_t b_lo = b_hi + (b_plane >> 4);
          uint32_t row = (uint32_t)(tap * p.dil);
          uint32_t d = tmem;
          for (int mt = 0; mt < p.MT; mt++, row += 128, d += (uint32_t)p.acc_stride) {
#pragma unroll
            for (int kk = 0; kk < TC_KB / 8; kk++) {
              const uint64_t ah = a_hi + row + kk * a_kk, al = a_lo + row + kk * a_kk;
              const uint64_t bh = b_hi + kk * b_kk, bl = b_lo + kk * b_kk;
              tc_mma_tf32(d, al, bh, idesc, (kb | tap | kk) ? 1u : 0u);   // small terms first
              tc_mma_tf32(d, ah, bl, idesc, 1u);
              tc_mma_tf32(d, ah, bh, idesc, 1u);
            }
          }
          tc_commit(&emptyB[sb]);                      // weight stage free once these MMAs retire
        }
        tc_commit(&emptyA[sa]);                        // slab free
      }
      tc_commit(accFull);
    }
  } else {
    // ===================== transform warps (activation + hi/lo split), then epilogue =====================
    const int wt = tid - 64;                           // 0..127
    const float* xb = p.x + (long long)b * p.xbs;
    const int rows_needed = p.MT * 128 + (p.taps - 1) * p.dil;
    const int items = rows_needed * TC_CHUNKS;
    const int t_first = t0 - p.pad;
    for (int kb = 0; kb < KBLOCKS; kb++) {
      const int sa = kb % TC_SA;
      mbar_wait(&emptyA[sa], ((kb / TC_SA) & 1) ^ 1);
      unsigned char* hi = a_base + (size_t)sa * 2 * a_plane;
      unsigned char* lo = hi + a_plane;
      const int c0 = kb * TC_KB;
      for (int base = 0; base < items; base += 128 * TC_LD) {
        float4 v[TC_LD];
#pragma unroll
        for (int u = 0; u < TC_LD; u++) {              // TC_LD independent 16-byte loads in flight per thread
          const int idx = base + u * 128 + wt;
          const int row = idx >> 2, ch = idx & 3;
          const int t = t_first + row;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx < items && t >= 0 && t < p.T) v[u] = __ldg(reinterpret_cast<const float4*>(xb + (long long)t * p.xrs + c0 + ch * 4));
        }
#pragma unroll
        for (int u = 0; u < TC_LD; u++) {
          const int idx = base + u * 128 + wt;
          if (idx >= items) continue;
          const int row = idx >> 2, ch = idx & 3;
          float4 a = v[u];
          if (p.in_act == FS2_ACT_LRELU) {
            a.x = a.x > 0.f ? a.x : a.x * p.in_slope; a.y = a.y > 0.f ? a.y : a.y * p.in_slope;
            a.z = a.z > 0.f ? a.z : a.z * p.in_slope; a.w = a.w > 0.f ? a.w : a.w * p.in_slope;
          }
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(a.x) & 0xffffe000u); l.x = a.x - h.x;
          h.y = __uint_as_float(__float_as_uint(a.y) & 0xffffe000u); l.y = a.y - h.y;
          h.z = __uint_as_float(__float_as_uint(a.z) & 0xffffe000u); l.z = a.z - h.z;
          h.w = __uint_as_float(__float_as_uint(a.w) & 0xffffe000u); l.w = a.w - h.w;
          const size_t off = ((size_t)ch * R + row) * 16;
          *reinterpret_cast<float4*>(hi + off) = h;
          *reinterpret_cast<float4*>(lo + off) = l;
        }
      }
      fence_proxy_async();                             // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&fullA[sa]);
    }

    // ---- epilogue: TMEM -> registers -> (per-warp smem transpose) -> coalesced global I/O ----
    // tcgen05.ld hands each thread one output ROW; writing rows straight out would touch 32 different 128-byte lines per
    // instruction.  Each warp therefore transposes its 32 x 32 block through a private 32 x 36 float staging tile (the slab
    // buffers are free once accFull has fired) so that 8 lanes cover one row's 128 bytes: every global load (residual,
    // accumulate) and store is a full-line access.
    mbar_wait(accFull, 0);
    tc_fence_after();
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    const int len_b = p.row_lens ? p.row_lens[b] : p.T;
    const int n0 = nblk * NB;
    float* stage = reinterpret_cast<float*>(a_base) + (warp - 2) * (32 * 36);
    for (int mt = 0; mt < p.MT; mt++) {
      const int row_base = t0 + mt * 128 + q * 32;
      for (int c = 0; c < NB; c += 32) {
        const int w = (NB - c) >= 32 ? 32 : 16;        // NB % 16 == 0
        {
          uint32_t v[32];
          const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * p.acc_stride + c);
          if (w == 32) {
            tc_ld32(taddr, v);
          } else {
            uint32_t v16[16];
            tc_ld16(taddr, v16);
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = v16[j];
          }
#pragma unroll
          for (int j = 0; j < 8; j++)
            if (j * 4 < w)
              *reinterpret_cast<float4*>(stage + lane * 36 + j * 4) =
                  make_float4(__uint_as_float(v[j * 4]), __uint_as_float(v[j * 4 + 1]), __uint_as_float(v[j * 4 + 2]),
                              __uint_as_float(v[j * 4 + 3]));
        }
        __syncwarp();
        const int lpr = w >> 2;                        // lanes per row (8 or 4)
        const int rpi = 32 / lpr;                      // rows per iteration (4 or 8)
        const int iters = 32 / rpi;                    // 8 or 4
        const int rr = lane / lpr, cc = (lane % lpr) * 4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + cc));
        float4 rv[8], yv[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {                  // all global loads of this block in flight before any store
          rv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          yv[k] = rv[k];
          const int t = row_base + k * rpi + rr;
          if (k < iters && t < p.T) {
            const long long off = (long long)t;
            if (p.res) rv[k] = *reinterpret_cast<const float4*>(p.res + (long long)b * p.rbs + off * p.rrs + n0 + c + cc);
            if (p.accumulate) yv[k] = *reinterpret_cast<const float4*>(p.y + (long long)b * p.ybs + off * p.yrs + n0 + c + cc);
          }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int r = k * rpi + rr;
          const int t = row_base + r;
          if (k < iters && t < p.T) {
            const float4 a = *reinterpret_cast<const float4*>(stage + r * 36 + cc);
            float4 o;
            o.x = (apply_act(a.x + bv.x, p.out_act, p.out_slope) + rv[k].x) * p.alpha + yv[k].x;
            o.y = (apply_act(a.y + bv.y, p.out_act, p.out_slope) + rv[k].y) * p.alpha + yv[k].y;
            o.z = (apply_act(a.z + bv.z, p.out_act, p.out_slope) + rv[k].z) * p.alpha + yv[k].z;
            o.w = (apply_act(a.w + bv.w, p.out_act, p.out_slope) + rv[k].w) * p.alpha + yv[k].w;
            if (t >= len_b) o = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(p.y + (long long)b * p.ybs + (long long)t * p.yrs + n0 + c + cc) = o;
          }
        }
        __syncwarp();                                  // staging tile is rewritten by the next block
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols));
  }
}

// ------------------------------------------------------------------ host side
static int pow2_cols(int c) {
  int v = 32;
  while (v < c) v <<= 1;
  return v;
}

int conv_tc_nb(int N) {  // output channels per CTA: at most 128 so that two CTAs (256 TMEM columns each) share an SM
  if (N % 16) return 0;
  if (N <= 128) return N;
  for (int nb = 128; nb >= 16; nb -= 16)
    if (N % nb == 0) return nb;
  return 0;
}

bool conv_tc_supported(const fs2_conv1d_args* a) {
  if (!a || a->Cin % TC_KB || a->N % 16 || conv_tc_nb(a->N) == 0) return false;
  if ((a->x_row_stride & 3) || (a->x_batch_stride & 3) || (a->y_row_stride & 3) || (a->y_batch_stride & 3)) return false;
  if (a->res && ((a->res_row_stride & 3) || (a->res_batch_stride & 3))) return false;
  if (a->in_act != FS2_ACT_NONE && a->in_act != FS2_ACT_LRELU) return false;
  if ((a->taps - 1) * a->dilation > 160) return false;
  return true;
}

// `wt` must be the tiled layout produced by fastspeech2_b200.packing.pack_conv_tc (see fs2b200.h)
int conv1d_tc(const fs2_conv1d_args* a, const float* wt, unsigned variant, cudaStream_t s) {
  if (!a || !a->x || !wt || !a->y) return FS2_ERR_ARG;
  if (a->B <= 0 || a->T <= 0 || a->Cin <= 0 || a->N <= 0 || a->taps <= 0) return FS2_ERR_ARG;
  if (!conv_tc_supported(a)) return FS2_ERR_UNSUPPORTED;
  if (!aligned16(a->x) || !aligned16(wt) || !aligned16(a->y) || (a->res && !aligned16(a->res))) return FS2_ERR_ARG;
  TcP p;
  p.x = a->x; p.xbs = a->x_batch_stride; p.xrs = a->x_row_stride;
  p.B = a->B; p.T = a->T; p.Cin = a->Cin;
  p.wt = wt; p.bias = a->bias; p.N = a->N; p.NB = conv_tc_nb(a->N);
  p.taps = a->taps; p.dil = a->dilation; p.pad = a->pad_left;
  p.in_act = a->in_act; p.in_slope = a->in_slope; p.out_act = a->out_act; p.out_slope = a->out_slope;
  p.res = a->res; p.rbs = a->res_batch_stride; p.rrs = a->res_row_stride;
  p.alpha = a->alpha; p.accumulate = a->accumulate; p.row_lens = a->row_lens;
  p.y = a->y; p.ybs = a->y_batch_stride; p.yrs = a->y_row_stride;
  p.variant = variant;
  p.acc_stride = (p.NB + 31) & ~31;
  const int halo = (a->taps - 1) * a->dilation;
  const int tiles128 = (a->T + 127) / 128;
  // Two CTAs per SM (one's epilogue / slab load overlaps the other's MMAs): aim for <= ~112 KB of shared memory and <= 256
  // TMEM columns per CTA; fall back to one CTA per SM when the halo makes the slab too large.
  const size_t bar_bytes = (2 * TC_SA + 2 * TC_SB_MAX + 1) * 8 + 16;
  const size_t budget2 = 112 * 1024, budget1 = 226 * 1024;
  int mt = 256 / p.acc_stride;
  if (mt > 2) mt = 2;
  if (mt > tiles128) mt = tiles128;
  if (mt < 1) mt = 1;
  size_t smem = 0;
  int sb = 0;
  for (;; mt--) {
    int R = mt * 128 + halo;
    R += (10 - (R & 7)) & 7;                           // R % 8 == 2: conflict-free transform stores
    p.R = R;
    const size_t a_bytes = (size_t)TC_SA * 2 * TC_CHUNKS * R * 16, b_stage = (size_t)2 * TC_CHUNKS * p.NB * 16;
    for (sb = TC_SB_MAX; sb >= 2; sb--) {
      smem = a_bytes + sb * b_stage + bar_bytes;
      if (smem <= budget2) break;
    }
    if (sb >= 2) break;
    if (mt == 1) {                                     // cannot fit two per SM: take what one CTA can have
      for (sb = TC_SB_MAX; sb >= 2; sb--) {
        smem = a_bytes + sb * b_stage + bar_bytes;
        if (smem <= budget1) break;
      }
      if (sb < 2) return FS2_ERR_UNSUPPORTED;
      break;
    }
  }
  p.MT = mt;
  p.SB = sb;
  p.tmem_cols = pow2_cols(mt * p.acc_stride);
  p.tiles_per_batch = (a->T + mt * 128 - 1) / (mt * 128);
  const long long gx = (long long)p.tiles_per_batch * a->B;
  if (gx > 0x7fffffffLL) return FS2_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return FS2_ERR_CUDA - (int)e;
    attr_set = true;
  }
  dim3 grid((unsigned)gx, a->N / p.NB);
  prof_before(s);
  conv_tc_kernel<<<grid, TC_THREADS, smem, s>>>(p);
  prof_after(s, 0, 2.0 * a->B * a->T * (double)a->Cin * a->taps * a->N);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

}  // namespace fs2
