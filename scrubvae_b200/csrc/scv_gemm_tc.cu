// tcgen05 / TMA overlapping-row GEMM and weight-gradient kernels (sm_100a), kind::tf32:
// fp32 operands as stored (the tensor core reads the TF32 part), fp32 accumulation in TMEM.
//
// Forward / dgrad (gemm_tc_kernel), one persistent CTA per SM, warp-specialised:
//   warp 0   TMA producer: A tile = 3-D box (32 k, bl rows of l, nb windows) of the halo-padded
//            channels-last activation viewed as (K, Lo, B) with OVERLAPPING strides (1, a_ls, a_bs) — the
//            implicit-GEMM im2col is done by the tensor map; W tile = 2-D box (32 k, bn) of the packed
//            [N][K] weights.  128-byte swizzle, out-of-bounds -> 0 (K tail, ragged l / b / n edges).
//   warp 1   one elected thread issues tcgen05.mma (M=128, N=bn<=256, K=8 per instruction), both operands
//            K-major from shared memory (descriptors built once, the low word advances per MMA).  Work item =
//            one 128-row tile with the accumulator double-buffered in TMEM (2 x 256 columns), or (sub = 2) two
//            row tiles against ONE W tile per stage, one accumulator each (less shared-memory traffic per FLOP).
//   warp 2   TMEM allocation.
//   warps 4-7 epilogue: tcgen05.ld -> swizzled float4 transpose through shared memory -> out_scale, bias,
//            residual, BatchNorm column sums, activation, TF32 rounding, float4 row stores; the arithmetic is
//            compiled per (activation, residual, sums, rounding) variant (epi_rows).
// gemm_tc2_kernel is the same body for CTA pairs (cta_group::2), kept as an experiment (SCV_TC_PAIR=1).
// Weight gradient (wgrad_tc_kernel): D[n][k] = sum_m dY[m][n] A[m][k]; both operands are MN-major in
// shared memory (dY rows are n-contiguous, A rows k-contiguous; 128B swizzle with 32-byte atoms, the only
// layout tcgen05 takes for MN-major tf32), the reduction runs over the rows of
// the same (32, bl, nb) boxes; work items (n tile, k tile, row split) are spread over the SMs and the
// epilogue adds into dW with coalesced red.global.add.v4.f32.  The bias gradient (column sums of dY) rides
// along as one extra N=16 MMA per 8 rows against a shared-memory tile of ones (k tile 0 items only).
#include <atomic>

#include "scv_tc.cuh"

namespace scv {
namespace tc {

encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (encode_tiled_fn)p;
  }
  return fn;
}

int make_tmap(CUtensorMap* tm, const void* base, int rank, const int64_t* dims, const int64_t* strides_elems,
              const int* box, const char* what, bool atom32, bool bf16) {
  encode_tiled_fn fn = get_encode_tiled();
  if (!fn) {
    set_error("%s: cuTensorMapEncodeTiled is not available from the driver", what);
    return -2;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = (cuuint64_t)strides_elems[i] * (bf16 ? 2 : sizeof(float));
  }
  CUresult r = fn(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d; dims %lld,%lld,%lld strides %lld,%lld box %d,%d,%d)",
              what, (int)r, (long long)dims[0], (long long)(rank > 1 ? dims[1] : 0), (long long)(rank > 2 ? dims[2] : 0),
              (long long)(rank > 1 ? strides_elems[1] : 0), (long long)(rank > 2 ? strides_elems[2] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return -3;
  }
  return 0;
}

}  // namespace tc
}  // namespace scv

namespace {

using namespace scv::tc;

constexpr int kEpiWarps = 8;        // two epilogue warps per TMEM lane quarter (they split the 32-column chunks): with one
                                    // warp per scheduler every instruction's latency was exposed (~2000 cycles per chunk)
constexpr int kThreads = 128 + 32 * kEpiWarps;  // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4.. epilogue
constexpr int kBM = 128;            // UMMA M (TMEM lanes)
constexpr int kBK = 32;             // k floats per stage row = one 128-byte swizzle span (TF32); bf16: 64 elements
constexpr int kRowBytes = 128;      // operand row of a stage, either element type
constexpr int kMaxBN = 256;         // UMMA N limit (cta_group::1)
constexpr int kTmemCols = 512;      // two accumulator buffers of 256 columns
constexpr int kXposeFloats = 32 * 32;  // per-warp epilogue transpose tile: 32 rows x 32 floats, 16-byte chunks XOR-swizzled
constexpr int kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA
constexpr int kMaxStages = 8;

struct GemmTcParams {
  int64_t B, Lo;
  int K, N;
  int bl, nb, lt, bt;  // A box rows = bl (l) x nb (windows); lt x bt row tiles
  int bn, n_tiles, m_tiles, k_chunks, stages;
  const float* bias;
  int bias_mod, bias_n;
  float* Y;
  int64_t y_bs, y_ls;
  int n_last;
  const float* R;
  int64_t r_bs, r_ls;
  int act, round_out;
  float out_scale;
  double* stats;
  int sub;           // row tiles per work item (1 or 2): with 2 the CTA runs two 128-row tiles against ONE W tile per
                     // stage (two accumulators, no TMEM double buffering) - a third less shared-memory traffic per FLOP
  int epi_kind;      // which compiled epilogue variant serves (act, R, stats, round_out); see gemm_tc_body
  // BatchNorm / PReLU backward reduction fused into a data-gradient epilogue (bnr_sums != nullptr): see scv_gemm_t
  const float* bnr_x;
  int64_t bnr_bs, bnr_ls;
  const float* bnr_chan;
  const float* bnr_slope;
  int bnr_c;
  double* bnr_sums;
  int ksplit;        // split-K: the reduction is cut into `ksplit` chunk ranges, one work item each, and the epilogue ADDS
                     // (red.global.add) into a pre-zeroed Y - for GEMMs with few output tiles and a long K (SCV_ACT_ACCUM)
  int kc_per;        // k chunks per split
  long long* trace;  // debug (SCV_TC_TRACE=<n events>): CTA 0 appends (role, event, tile, clock) records
  int* work;         // dynamic work distribution (nullptr: static round robin): work[0] = next item, work[1] = CTAs done.
                     // CTAs that start late (SMs held by a concurrent NCCL kernel) or draw long items no longer decide the
                     // launch's duration; the last CTA to leave resets both counters, so launches and graph replays are
                     // self-cleaning
};

// debug timeline of CTA 0: role 0 producer / 1 mma / 2 epilogue warp 0; only active when p.trace != nullptr
__device__ __forceinline__ void tc_trace(const GemmTcParams& p, int role, int ev, int tile) {
  if (p.trace && blockIdx.x == 0) {
    const unsigned long long i = atomicAdd(reinterpret_cast<unsigned long long*>(p.trace), 1ULL);
    if (i < 4000) {
      p.trace[1 + 2 * i] = ((long long)role << 40) | ((long long)ev << 32) | (unsigned)tile;
      p.trace[2 + 2 * i] = clock64();
    }
  }
}

struct SmemCtl {  // lives after the operand stages
  uint64_t full[kMaxStages], empty[kMaxStages], tfull[2], tempty[2];
  uint64_t sfull[2], sempty[2];  // dynamic scheduler: the producer publishes item indices through item[2]
  int item[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// consumer side of the dynamic scheduler (MMA thread, epilogue warps): next item index, -1 = no more work
struct ItemReader {
  uint32_t sfull, sempty;  // shared-memory addresses of slot 0's barriers (slot 1: + 8)
  const int* item;
  int slot;
  uint32_t phase;
  template <typename Ctl>
  __device__ __forceinline__ void init(Ctl* ctl) {
    sfull = smem_u32(&ctl->sfull[0]);
    sempty = smem_u32(&ctl->sempty[0]);
    item = ctl->item; slot = 0; phase = 0;
  }
  // whole-warp callers pass warp_wide = true: every lane reads, one lane hands the slot back
  __device__ __forceinline__ int next(bool warp_wide, int lane) {
    mbar_wait(sfull + 8u * slot, phase);
    const int t = *reinterpret_cast<const volatile int*>(item + slot);
    if (warp_wide) {
      __syncwarp();
      if (lane == 0) mbar_arrive(sempty + 8u * slot);
    } else {
      mbar_arrive(sempty + 8u * slot);
    }
    if (++slot == 2) { slot = 0; phase ^= 1; }
    return t;
  }
};
// producer side: publish item t (or -1)
struct ItemWriter {
  uint32_t sfull, sempty;
  int* item;
  int slot;
  uint32_t phase;
  template <typename Ctl>
  __device__ __forceinline__ void init(Ctl* ctl) {
    sfull = smem_u32(&ctl->sfull[0]);
    sempty = smem_u32(&ctl->sempty[0]);
    item = ctl->item; slot = 0; phase = 0;
  }
  __device__ __forceinline__ void put(int t) {
    mbar_wait(sempty + 8u * slot, phase ^ 1);
    *reinterpret_cast<volatile int*>(item + slot) = t;
    mbar_arrive(sfull + 8u * slot);  // release: the store above is visible to whoever acquires the phase
    if (++slot == 2) { slot = 0; phase ^= 1; }
  }
};
// leaving a dynamically scheduled launch (called by the producer thread once it has drawn its sentinel): the last CTA
// resets the counters - every CTA has made its last draw by then, and kernel boundaries order the reset against the
// next launch that uses the slot
__device__ __forceinline__ void work_leave(int* work) {
  const int d = atomicAdd(work + 1, 1);
  if (d == (int)gridDim.x - 1) {
    work[0] = 0;
    work[1] = 0;
    __threadfence();
  }
}
// BatchNorm column sums / sums of squares of this CTA's tiles (flushed when the n tile changes): 2 x kMaxBN floats placed
// behind the transpose tiles ONLY for launches that collect statistics — 2 KB that decide whether a third 64 KB stage
// fits for the 256 x 256 work items.

// epilogue warps only (threads 128..255): named barrier 1
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); }
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float act_apply(float v, int act, float r) {
  if (act == SCV_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == SCV_ACT_TANH) return tanhf(v);
  if (act == SCV_ACT_RELUMASK) return r > 0.f ? v : 0.f;
  return v;
}

// Epilogue arithmetic of one 32x32 chunk for lane (rq, cq): rows 4 i + rq (i < 8), columns n .. n + 3.
// ACT >= 0 / RD / ST / RND are compile-time; ACT = -1 is the generic path (activation, rounding from p at run time).
// BNR: the stored value is the gradient w.r.t. the OUTPUT of a BatchNorm(+PReLU) layer whose pre-normalisation input X has
// the same (row, column) addressing; its backward reduction (sum g', sum g' xhat per channel, PReLU slope gradient) is
// accumulated here instead of in a separate pass over X and Y (s1 / s2 carry the two column sums, ds the slope term).
template <int ACT, bool RD, bool ST, bool RND, bool ACC = false, bool BNR = false>
__device__ __forceinline__ void epi_rows(const GemmTcParams& p, const float4* xp4, const int rq, const int cq, const bool col_ok,
                                         const int n, const float4 bv, const long long* __restrict__ yoff,
                                         const long long* __restrict__ roff, const int* __restrict__ ncap, float4& s1,
                                         float4& s2, const long long* __restrict__ xoff = nullptr,
                                         float* ds = nullptr) {
  const int act = ACT >= 0 ? ACT : p.act;
  const bool rnd = ACT >= 0 ? RND : (p.round_out != 0);
  const float osc = p.out_scale;
  float4 c_sc = make_float4(1.f, 1.f, 1.f, 1.f), c_sh = make_float4(0.f, 0.f, 0.f, 0.f), c_mu = c_sh, c_rs = c_sc;
  float slope = 0.f;
  bool has_act = false;
  if (BNR) {
    if (p.bnr_chan && col_ok) {
      const int ch = n % p.bnr_c, C = p.bnr_c;
      c_sc = __ldg(reinterpret_cast<const float4*>(p.bnr_chan + ch));
      c_sh = __ldg(reinterpret_cast<const float4*>(p.bnr_chan + C + ch));
      c_mu = __ldg(reinterpret_cast<const float4*>(p.bnr_chan + 2 * C + ch));
      c_rs = __ldg(reinterpret_cast<const float4*>(p.bnr_chan + 3 * C + ch));
    }
    has_act = p.bnr_slope != nullptr;
    if (has_act) slope = __ldg(p.bnr_slope);
  }
  // global loads (residual R, BatchNorm input X) are issued G rows deep before the arithmetic; the BNR variants hold two
  // operands per row and go 4 deep to stay inside the register budget
  constexpr int G = BNR ? 4 : 8;
#pragma unroll
  for (int i0 = 0; i0 < 8; i0 += G) {
    float4 rv[RD ? G : 1], xv[BNR ? G : 1];
    if (RD) {
#pragma unroll
      for (int u = 0; u < G; ++u) {
        rv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok && n < ncap[i0 + u]) rv[u] = __ldg(reinterpret_cast<const float4*>(p.R + roff[i0 + u] + n));
      }
    }
    if (BNR) {
#pragma unroll
      for (int u = 0; u < G; ++u) {
        xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok && n < ncap[i0 + u]) xv[u] = __ldg(reinterpret_cast<const float4*>(p.bnr_x + xoff[i0 + u] + n));
      }
    }
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const int i = i0 + u;
      const int r = i * 4 + rq;
      const float4 a = xp4[r * 8 + (cq ^ (r & 7))];
      if (col_ok && n < ncap[i]) {
        float4 tv = make_float4(fmaf(osc, a.x, bv.x), fmaf(osc, a.y, bv.y), fmaf(osc, a.z, bv.z), fmaf(osc, a.w, bv.w));
        float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (RD) {
          r4 = rv[u];
          if (act != SCV_ACT_RELUMASK) { tv.x += r4.x; tv.y += r4.y; tv.z += r4.z; tv.w += r4.w; }
        }
        if (ST) {
          s1.x += tv.x; s1.y += tv.y; s1.z += tv.z; s1.w += tv.w;
          s2.x = fmaf(tv.x, tv.x, s2.x); s2.y = fmaf(tv.y, tv.y, s2.y); s2.z = fmaf(tv.z, tv.z, s2.z); s2.w = fmaf(tv.w, tv.w, s2.w);
        }
        if (BNR) {
          const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, gs[4] = {tv.x, tv.y, tv.z, tv.w};
          const float sc[4] = {c_sc.x, c_sc.y, c_sc.z, c_sc.w}, sh[4] = {c_sh.x, c_sh.y, c_sh.z, c_sh.w};
          const float mu[4] = {c_mu.x, c_mu.y, c_mu.z, c_mu.w}, rs[4] = {c_rs.x, c_rs.y, c_rs.z, c_rs.w};
          float a1[4], a2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float v = fmaf(xs[q], sc[q], sh[q]);
            float gg = gs[q];
            if (has_act && v < 0.f) { *ds += gg * v; gg *= slope; }
            a1[q] = gg;
            a2[q] = gg * (xs[q] - mu[q]) * rs[q];
          }
          s1.x += a1[0]; s1.y += a1[1]; s1.z += a1[2]; s1.w += a1[3];
          s2.x += a2[0]; s2.y += a2[1]; s2.z += a2[2]; s2.w += a2[3];
        }
        float4 yv = tv;
        if (act != SCV_ACT_NONE)
          yv = make_float4(act_apply(tv.x, act, r4.x), act_apply(tv.y, act, r4.y), act_apply(tv.z, act, r4.z),
                           act_apply(tv.w, act, r4.w));
        if (rnd) yv = scv::round_tf32(yv);
        if (ACC) red_add_v4(p.Y + yoff[i] + n, yv);
        else *reinterpret_cast<float4*>(p.Y + yoff[i] + n) = yv;
      }
    }
  }
}

// kCtas = 1: one CTA per 128-row tile (cta_group::1).  kCtas = 2: the two CTAs of a cluster share a 256-row tile
// (cta_group::2): each loads its own 128 A rows and HALF of the W tile, the leader's single thread issues M = 256
// MMAs that read W from both shared memories — per SM the shared-memory fill and operand-read traffic per FLOP
// drops by a third, which is what paces the large TF32 layers.
// kMc (kCtas = 1 only): the two CTAs of a cluster work on different row tiles of the SAME n tile in lock step; each loads
// half of the W tile and MULTICASTS it into both shared memories, so W crosses the L2 -> SM fabric once per pair.  The
// layers of this network are paced by that fabric (12.3 TB/s chip-wide, ~42 B/cycle per SM, against the 64-96 B/cycle a
// 128x256 TF32 tile consumes), not by the tensor pipe.
// kBnr: the epilogue variants with the fused BatchNorm backward reduction live in their own kernel instantiation — their
// extra per-row offsets and per-channel constants cost registers (the combined kernel spilled 200 bytes and every GEMM of
// the step slowed down by 8 %).
template <int kCtas, bool kMc = false, bool kBf16 = false, bool kBnr = false>
__device__ __forceinline__ void gemm_tc_body(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmTcParams& p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic shared memory is only guaranteed 16-byte aligned: round up to the 1024 B the swizzle needs
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kCtas == 2 ? cluster_ctarank() : 0u;
  const uint32_t crank = kMc ? cluster_ctarank() : 0u;
  constexpr int kGroup = kMc ? 2 : kCtas;  // CTAs walking one work-item list together
  const int bnw = p.bn / kCtas;                       // W rows held by this CTA
  const int sub = kCtas == 1 ? p.sub : 1;
  constexpr int kEK = kBf16 ? 64 : 32;                // operand elements per 128-byte stage row
  const uint32_t a_tile = kBM * kRowBytes;            // 16 KB per 128-row tile
  const uint32_t a_bytes = a_tile * (uint32_t)sub;
  const uint32_t w_bytes = (uint32_t)bnw * kRowBytes; // rows of 128 B (multiple of 2 KB)
  const uint32_t stage_bytes = a_bytes + w_bytes;
  uint8_t* ctl_raw = smem + (size_t)p.stages * stage_bytes;
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(ctl_raw);
  float* xpose = reinterpret_cast<float*>(ctl_raw + ((sizeof(SmemCtl) + 15) & ~size_t(15)));  // kEpiWarps x 32 x 32 floats
  float* sacc = xpose + kEpiWarps * kXposeFloats;  // [2][kMaxBN], present only when p.stats != nullptr
  const uint32_t tile_tx = ((uint32_t)(p.bl * p.nb) * kRowBytes * (uint32_t)sub + w_bytes) * kCtas;
  // work items: (n tile, group of kCtas * sub consecutive m tiles); the CTAs of a pair walk the same list
  const int mper = kGroup * sub;
  const int m_groups = (p.m_tiles + mper - 1) / mper;
  const int tiles = p.n_tiles * m_groups;
  const int total = tiles * p.ksplit;  // item t: k split t / tiles, then (n tile, m group)
  const int first = (int)blockIdx.x / kGroup, step = (int)gridDim.x / kGroup;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      // pair: ONE arrival (the leader's expect_tx for both CTAs' bytes).  Round 1 also had the peer arrive on the leader's
      // barrier every chunk: that mbarrier.arrive.release.cluster stalled the peer's producer ~1500 cycles per stage
      // and made the pair kernel 2x slower than one CTA (tools/cu/pair_pipe.cu: the pipeline itself runs at the MMA rate)
      mbar_init(smem_u32(&ctl->full[s]), 1);
      mbar_init(smem_u32(&ctl->empty[s]), kMc ? 2 : 1);  // multicast pair: both CTAs' MMAs must be done with the stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&ctl->tfull[i]), 1);
      mbar_init(smem_u32(&ctl->tempty[i]), kEpiWarps * kCtas);
      mbar_init(smem_u32(&ctl->sfull[i]), 1);
      mbar_init(smem_u32(&ctl->sempty[i]), 1 + kEpiWarps);
    }
    fence_barrier_init();
  }
  // dynamic work distribution (single-CTA kernels): every CTA starts on item blockIdx.x and DRAWS the following ones
  // (gridDim.x + counter).  A draw for the first item too was measured slower: 148 same-address atomics at launch
  // serialise in L2 (~2 us for the last CTA), while later draws are staggered and hidden behind the item's loads.
  const bool dyn = kCtas == 1 && !kMc && p.work != nullptr;
  const int t_first = first;
  if (warp == 2) {
    if (kCtas == 2) tmem_alloc2(smem_u32(&ctl->tmem_base), kTmemCols);
    else tmem_alloc(smem_u32(&ctl->tmem_base), kTmemCols);
  }
  tc_fence_before();
  if (kCtas == 2 || kMc) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      ItemWriter iw;
      iw.init(ctl);
      for (int t = t_first;;) {
        if (dyn) iw.put(t < total ? t : -1);
        if (t >= total) {
          if (dyn) work_leave(p.work);  // hidden behind the last item's MMAs and epilogue
          break;
        }
        const int t_next = dyn ? step + atomicAdd(p.work, 1) : t + step;  // drawn now, needed after this item's loads
        const int ks = t / tiles, tt = t - ks * tiles;
        const int nt = tt / m_groups, mt = (tt - nt * m_groups) * mper + (kMc ? (int)crank * sub : (int)rank);
        const int bt_i = mt / p.lt, lt_i = mt - bt_i * p.lt;  // mt >= m_tiles (odd tail): box fully out of bounds -> zeros
        const int l0 = lt_i * p.bl, b0 = bt_i * p.nb, n0 = nt * p.bn + (int)rank * bnw;
        const int bt_2 = (mt + 1) / p.lt, lt_2 = (mt + 1) - bt_2 * p.lt;  // second row tile (sub == 2)
        const int kc0 = ks * p.kc_per, kc1 = min(p.k_chunks, kc0 + p.kc_per);
        tc_trace(p, 0, 0, t);
        for (int kc = kc0; kc < kc1; ++kc) {
          mbar_wait(smem_u32(&ctl->empty[s]), ph ^ 1);
          if (kc == kc0 || kc == kc1 - 1) tc_trace(p, 0, 1 + (kc != kc0), t);
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          if (kCtas == 2) {
            const uint32_t fb = mapa_rank(smem_u32(&ctl->full[s]), 0);  // the leader's barrier counts both CTAs' bytes
            if (rank == 0) mbar_expect_tx(smem_u32(&ctl->full[s]), tile_tx);
            tma2_load_3d(sa, &tmA, fb, kc * kEK, l0, b0);
            tma2_load_2d(sa + a_bytes, &tmW, fb, kc * kEK, n0);
          } else {
            const uint32_t fb = smem_u32(&ctl->full[s]);
            mbar_expect_tx(fb, tile_tx);
            tma_load_3d(sa, &tmA, fb, kc * kEK, l0, b0);
            if (sub == 2) tma_load_3d(sa + a_tile, &tmA, fb, kc * kEK, lt_2 * p.bl, bt_2 * p.nb);
            if (kMc)  // my half of the W tile, into both CTAs' stage
              tma_load_2d_mc(sa + a_bytes + crank * (w_bytes >> 1), &tmW, fb, kc * kEK, n0 + (int)crank * (p.bn >> 1), (uint16_t)3);
            else
              tma_load_2d(sa + a_bytes, &tmW, fb, kc * kEK, n0);
          }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        t = t_next;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = kBf16 ? idesc_bf16(kBM, p.bn, 0, 0) : idesc_tf32(kBM * kCtas, p.bn, 0, 0);
      // K-major, 128B swizzle: LBO 16 B (unused), SBO 1024 B; the stage base addresses are 1024-byte aligned
      const uint64_t desc0 = smem_desc(smem_u32(smem), 16, 1024);
      const uint32_t desc_lo0 = (uint32_t)desc0, desc_hi = (uint32_t)(desc0 >> 32);
      const uint32_t stage_units = stage_bytes >> 4;
      int s = 0, it = 0;
      uint32_t ph = 0;
      ItemReader ir;
      ir.init(ctl);
      for (int t = first;; ++it) {
        if (dyn) t = ir.next(false, 0); else if (it) t += step;
        if (t < 0 || t >= total) break;
        const int acc = sub == 2 ? 0 : (it & 1);  // sub == 2: both TMEM halves hold this item's two accumulators
        tc_trace(p, 1, 0, t);
        mbar_wait(smem_u32(&ctl->tempty[acc]), (sub == 2 ? (it & 1) : ((it >> 1) & 1)) ^ 1);
        tc_fence_after();
        tc_trace(p, 1, 1, t);
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kMaxBN;
        const int ks = t / tiles;
        const int kc0 = ks * p.kc_per, kc1 = min(p.k_chunks, kc0 + p.kc_per);
        for (int kc = kc0; kc < kc1; ++kc) {
          mbar_wait(smem_u32(&ctl->full[s]), ph);
          tc_fence_after();
          if (kc == kc0 || kc == kc1 - 1) tc_trace(p, 1, 2 + (kc != kc0), t);
          // descriptor low words of this stage; each k step of 8 floats advances the start address by 32 B = 2 units
          const uint32_t a_lo = desc_lo0 + (uint32_t)s * stage_units, b_lo = a_lo + (a_bytes >> 4);
          const uint32_t acc0 = kc != kc0 ? 1u : 0u;
          if (kCtas == 2) {
            umma2_tf32_lh(d_tmem, a_lo, desc_hi, b_lo, desc_hi, idesc, acc0);
            umma2_tf32_lh(d_tmem, a_lo + 2, desc_hi, b_lo + 2, desc_hi, idesc, 1u);
            umma2_tf32_lh(d_tmem, a_lo + 4, desc_hi, b_lo + 4, desc_hi, idesc, 1u);
            umma2_tf32_lh(d_tmem, a_lo + 6, desc_hi, b_lo + 6, desc_hi, idesc, 1u);
          } else {
            // 4 k steps of 32 bytes (8 tf32 / 16 bf16 elements) inside the 128-byte swizzle span
            umma_lh<kBf16>(d_tmem, a_lo, desc_hi, b_lo, desc_hi, idesc, acc0);
            umma_lh<kBf16>(d_tmem, a_lo + 2, desc_hi, b_lo + 2, desc_hi, idesc, 1u);
            umma_lh<kBf16>(d_tmem, a_lo + 4, desc_hi, b_lo + 4, desc_hi, idesc, 1u);
            umma_lh<kBf16>(d_tmem, a_lo + 6, desc_hi, b_lo + 6, desc_hi, idesc, 1u);
            if (sub == 2) {
              const uint32_t a2 = a_lo + (a_tile >> 4), d2 = d_tmem + kMaxBN;
              umma_lh<kBf16>(d2, a2, desc_hi, b_lo, desc_hi, idesc, acc0);
              umma_lh<kBf16>(d2, a2 + 2, desc_hi, b_lo + 2, desc_hi, idesc, 1u);
              umma_lh<kBf16>(d2, a2 + 4, desc_hi, b_lo + 4, desc_hi, idesc, 1u);
              umma_lh<kBf16>(d2, a2 + 6, desc_hi, b_lo + 6, desc_hi, idesc, 1u);
            }
          }
          if (kCtas == 2) umma2_commit(smem_u32(&ctl->empty[s]));
          else if (kMc) umma_commit_mc(smem_u32(&ctl->empty[s]), (uint16_t)3);
          else umma_commit(smem_u32(&ctl->empty[s]));
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (kCtas == 2) umma2_commit(smem_u32(&ctl->tfull[acc])); else umma_commit(smem_u32(&ctl->tfull[acc]));
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // Epilogue.  Warp ew owns TMEM lanes [32 ew, 32 ew + 32) = tile rows.  Per 32-column chunk: tcgen05.ld (thread =
    // row) -> swizzled float4 transpose through shared memory -> lane (rq, cq) holds 4 consecutive columns of rows
    // 4 i + rq (i < 8): 128-byte contiguous row segments per 8 lanes for the residual loads and the stores.
    // warp -> TMEM lane quarter q (a warp can only read lanes 32 (warp % 4) .. +31) and column phase h: the kEpiWarps / 4
    // warps of a quarter take every (kEpiWarps / 4)-th 32-column chunk
    const int ew = warp - 4, q = ew & 3, h = ew >> 2;
    constexpr int kColStep = 32 * (kEpiWarps / 4);
    const int rq = lane >> 3, cq = lane & 7;
    float4* xp4 = reinterpret_cast<float4*>(xpose + ew * kXposeFloats);
    const int rows_in_box = p.bl * p.nb;
    const int N = p.N;
    const int et = threadIdx.x - 128;
    const bool has_bnr = kBnr && p.bnr_sums != nullptr;
    // column sums through the shared-memory accumulators (a PReLU-only reduction, bnr_chan == nullptr, has none)
    const bool has_stats = p.stats != nullptr || (has_bnr && p.bnr_chan != nullptr);
    // where the column sums go: forward statistics stats[n], stats[N + n]; backward reduction sums[n % C], sums[C + n % C]
    double* const sum_base = has_bnr ? p.bnr_sums : p.stats;
    const int sum_mod = has_bnr ? p.bnr_c : p.N, sum_off2 = has_bnr ? p.bnr_c : p.N;
    float ds_acc = 0.f;
    if (has_stats) {
      for (int c = et; c < 2 * kMaxBN; c += 32 * kEpiWarps) sacc[c] = 0.f;
      epi_bar();
    }
    const uint32_t tempty_addr[2] = {kCtas == 2 ? mapa_rank(smem_u32(&ctl->tempty[0]), 0) : smem_u32(&ctl->tempty[0]),
                                     kCtas == 2 ? mapa_rank(smem_u32(&ctl->tempty[1]), 0) : smem_u32(&ctl->tempty[1])};
    int it = 0, cur_nt = -1;
    ItemReader ir;
    ir.init(ctl);
    for (int t = first;; ++it) {
      if (dyn) t = ir.next(true, lane); else if (it) t += step;
      if (t < 0 || t >= total) break;
      const int ks = t / tiles, tt = t - ks * tiles;
      const int nt = tt / m_groups, mt0 = (tt - nt * m_groups) * mper + (kMc ? (int)crank * sub : (int)rank);
      const int n0 = nt * p.bn;
      const int ncols = min(p.bn, N - n0);
      if (has_stats && nt != cur_nt) {
        if (cur_nt >= 0) {  // flush the finished n tile's column sums
          epi_bar();
          const int pn0 = cur_nt * p.bn, pc = min(p.bn, N - pn0);
          for (int c = et; c < pc; c += 32 * kEpiWarps) {
            atomicAdd(sum_base + (pn0 + c) % sum_mod, (double)sacc[c]);
            atomicAdd(sum_base + sum_off2 + (pn0 + c) % sum_mod, (double)sacc[kMaxBN + c]);
            sacc[c] = 0.f;
            sacc[kMaxBN + c] = 0.f;
          }
          epi_bar();
        }
        cur_nt = nt;
      }
      const int acc = sub == 2 ? 0 : (it & 1);
      if (ew == 0 && lane == 0) tc_trace(p, 2, 0, t);
      mbar_wait(smem_u32(&ctl->tfull[acc]), sub == 2 ? (it & 1) : ((it >> 1) & 1));
      tc_fence_after();
      if (ew == 0 && lane == 0) tc_trace(p, 2, 1, t);
      bool released = false;
     for (int sj = 0; sj < sub; ++sj) {
      const int mt = mt0 + sj;
      const int bt_i = mt / p.lt, lt_i = mt - bt_i * p.lt;
      long long yoff[8], roff[kBnr ? 1 : 8];  // kBnr: X and R share Y's row geometry (checked on the host): one offset array
      int ncap[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = q * 32 + i * 4 + rq;
        const int bi = row / p.bl, li = row - bi * p.bl;
        const int64_t b = (int64_t)bt_i * p.nb + bi, l = (int64_t)lt_i * p.bl + li;
        const bool ok = row < rows_in_box && mt < p.m_tiles && b < p.B && l < p.Lo;
        yoff[i] = (long long)(b * p.y_bs + l * p.y_ls);
        if (!kBnr) roff[kBnr ? 0 : i] = (long long)(b * p.r_bs + l * p.r_ls);
        ncap[i] = ok ? ((l == p.Lo - 1) ? p.n_last : N) : 0;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub == 2 ? sj : acc) * kMaxBN;
      for (int c0 = h * 32; c0 < ncols; c0 += kColStep) {
        // the bias columns of this lane: loaded BEFORE the TMEM read so that the global-load latency hides behind it
        // (ncu: the first FFMA of every chunk sat ~500 cycles on this load)
        const int cc = c0 + cq * 4;
        const int n = n0 + cc;
        const bool col_ok = cc < ncols;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && col_ok && n < p.bias_n && ks == 0) bv = __ldg(reinterpret_cast<const float4*>(p.bias + (n % p.bias_mod)));
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        if (c0 + kColStep >= ncols && sj == sub - 1) {  // this warp's last read of the accumulator(s): hand the buffer back
          released = true;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kCtas == 2) mbar_arrive_cluster(tempty_addr[acc]); else mbar_arrive(tempty_addr[acc]);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          xp4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
        const bool rd = p.R != nullptr;
        // one instantiation per (activation, residual, sums, rounding) combination the step uses: the per-element
        // work is then a handful of instructions instead of a runtime-branching generic path (one epilogue warp
        // per scheduler: every instruction's latency is exposed)
        if (kBnr) {
          if (rd) epi_rows<SCV_ACT_NONE, true, false, false, false, true>(p, xp4, rq, cq, col_ok, n, bv, yoff, yoff, ncap, s1, s2, yoff, &ds_acc);
          else epi_rows<SCV_ACT_NONE, false, false, false, false, true>(p, xp4, rq, cq, col_ok, n, bv, yoff, yoff, ncap, s1, s2, yoff, &ds_acc);
        } else {
        switch (p.epi_kind) {
          case 0: epi_rows<SCV_ACT_NONE, false, false, false>(p, xp4, rq, cq, col_ok, n, bv, yoff, roff, ncap, s1, s2); break;
          case 1: epi_rows<SCV_ACT_NONE, false, true, false>(p, xp4, rq, cq, col_ok, n, bv, yoff, roff, ncap, s1, s2); break;
          case 2: epi_rows<SCV_ACT_NONE, true, false, false>(p, xp4, rq, cq, col_ok, n, bv, yoff, roff, ncap, s1, s2); break;
          case 3: epi_rows<SCV_ACT_NONE, true, true, false>(p, xp4, rq, cq, col_ok, n, bv, yoff, roff, ncap, s1, s2); break;
          case 4: epi_rows<SCV_ACT_NONE, false, false, true>(p, xp4, rq, cq, col_ok, n, bv, yoff, roff, ncap, s1, s2); break;
          case 5: epi_rows<SCV_ACT_TANH, false, false, false>(p, xp4, rq, cq, col_ok, n, bv, yoff, roff, ncap, s1, s2); break;
          case 6: epi_rows<SCV_ACT_NONE, false, false, false, true>(p, xp4, rq, cq, col_ok, n, bv, yoff, roff, ncap, s1, s2); break;
          default:
            if (rd) epi_rows<-1, true, true, true>(p, xp4, rq, cq, col_ok, n, bv, yoff, roff, ncap, s1, s2);
            else epi_rows<-1, false, true, true>(p, xp4, rq, cq, col_ok, n, bv, yoff, roff, ncap, s1, s2);
        }
        }
        if (has_stats) {
#pragma unroll
          for (int off = 8; off <= 16; off <<= 1) {
            s1.x += __shfl_xor_sync(0xffffffffu, s1.x, off); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, off);
            s1.z += __shfl_xor_sync(0xffffffffu, s1.z, off); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, off);
            s2.x += __shfl_xor_sync(0xffffffffu, s2.x, off); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, off);
            s2.z += __shfl_xor_sync(0xffffffffu, s2.z, off); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, off);
          }
          if (rq == 0 && col_ok) {
            atomicAdd(sacc + cc, s1.x); atomicAdd(sacc + cc + 1, s1.y);
            atomicAdd(sacc + cc + 2, s1.z); atomicAdd(sacc + cc + 3, s1.w);
            atomicAdd(sacc + kMaxBN + cc, s2.x); atomicAdd(sacc + kMaxBN + cc + 1, s2.y);
            atomicAdd(sacc + kMaxBN + cc + 2, s2.z); atomicAdd(sacc + kMaxBN + cc + 3, s2.w);
          }
        }
        __syncwarp();
      }
     }  // sub tiles
      if (!released) {  // no chunk of the last tile fell to this warp: it still owes its arrival
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kCtas == 2) mbar_arrive_cluster(tempty_addr[acc]); else mbar_arrive(tempty_addr[acc]);
        }
      }
      if (ew == 0 && lane == 0) tc_trace(p, 2, 2, t);
    }
    if (has_stats && cur_nt >= 0) {
      epi_bar();
      const int pn0 = cur_nt * p.bn, pc = min(p.bn, N - pn0);
      for (int c = et; c < pc; c += 32 * kEpiWarps) {
        atomicAdd(sum_base + (pn0 + c) % sum_mod, (double)sacc[c]);
        atomicAdd(sum_base + sum_off2 + (pn0 + c) % sum_mod, (double)sacc[kMaxBN + c]);
      }
    }
    if (has_bnr && p.bnr_slope != nullptr) {  // PReLU slope gradient: sum over everything this warp touched
      const float d = scv::warp_sum(ds_acc);
      if (lane == 0 && d != 0.f) atomicAdd(p.bnr_sums + 2 * p.bnr_c, (double)d);
    }
  }
  tc_fence_before();
  if (kCtas == 2 || kMc) cluster_sync_all(); else __syncthreads();  // pair: no CTA may exit while its peer still signals it
  if (warp == 2) {
    tc_fence_after();
    if (kCtas == 2) tmem_dealloc2(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmTcParams p) {
  gemm_tc_body<1>(tmA, tmW, p);
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_bnr_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmTcParams p) {
  gemm_tc_body<1, false, false, true>(tmA, tmW, p);
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_bnr_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmTcParams p) {
  gemm_tc_body<1, false, true, true>(tmA, tmW, p);
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmTcParams p) {
  gemm_tc_body<1, false, true>(tmA, tmW, p);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc_mc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmTcParams p) {
  gemm_tc_body<1, true>(tmA, tmW, p);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmTcParams p) {
  gemm_tc_body<2>(tmA, tmW, p);
}

// ---------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------
struct WgradTcParams {
  int K, N;
  int bl, nb, lt, bt;   // row boxes as above; R = bl*nb rows (multiple of 8) reduced per stage
  int bnk;              // k tile width (UMMA N), multiple of 16, <= 256
  int n_tiles, k_tiles, splits, groups, gps, stages;
  int sub;  // 128-row n tiles per work item (1 or 2).  2: both share the A slabs of a stage (two accumulators in the
            // two TMEM halves, no double buffering) - a third less L2 -> shared-memory traffic per FLOP
  float* dW;
  float* dbias;  // bias gradient = column sums of dY: one extra N=16 MMA per 8 rows against a tile of ones (k tile 0 only)
  int bias_mod, bias_n;
  long long* trace;  // debug (SCV_TC_TRACE): as GemmTcParams::trace
  int* work;         // dynamic work distribution, as GemmTcParams::work
};

__device__ __forceinline__ void tc_trace_w(const WgradTcParams& p, int role, int ev, int tile) {
  if (p.trace && blockIdx.x == 0) {
    const unsigned long long i = atomicAdd(reinterpret_cast<unsigned long long*>(p.trace), 1ULL);
    if (i < 4000) {
      p.trace[1 + 2 * i] = ((long long)role << 40) | ((long long)ev << 32) | (unsigned)tile;
      p.trace[2 + 2 * i] = clock64();
    }
  }
}

constexpr int kWgradMaxBNK = 224;  // k-tile width limit (whole 32-float slabs); 224 + 16 bias columns fit a 256-column TMEM buffer
constexpr int kBiasCol = 240;      // TMEM column of the bias accumulator inside each 256-column buffer

struct SmemCtlW {
  uint64_t full[kMaxStages], empty[kMaxStages], tfull[2], tempty[2];
  uint64_t sfull[2], sempty[2];
  int item[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// kMc: the two CTAs of a cluster take ADJACENT n tiles of the same (k tile, row split): they need the same A slabs, so each
// loads half of them and multicasts into both shared memories (dY slabs are private).  L2 -> SM traffic per FLOP drops by a
// third (30 KB instead of 44 KB per 32-row stage), which is what bounds this kernel.
// kBf16: bf16 operands — slabs are 64 elements wide (still 128 bytes per row), one MMA reduces 16 rows, plain 128-byte
// swizzle (layout type 2, 8-row atoms) instead of the 32-byte-atom variant fp32 MN-major operands need.
template <bool kMc, bool kBf16 = false>
__device__ __forceinline__ void wgrad_tc_body(const CUtensorMap& tmY, const CUtensorMap& tmA, const WgradTcParams& p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.bl * p.nb;
  constexpr int SW = kBf16 ? 64 : 32;       // slab width in operand elements (128 bytes)
  constexpr int kYSlabs = kBM / SW;         // dY slabs per 128-row n tile
  constexpr int kRowsPerMma = kBf16 ? 16 : 8;
  const uint32_t slab = (uint32_t)R * 128;  // one slab of R rows (R % 8 == 0 -> 1 KB multiple)
  const int a_slabs = (p.bnk + SW - 1) / SW;
  const int sub = p.sub;
  const uint32_t y_bytes = kYSlabs * slab * (uint32_t)sub, a_bytes = (uint32_t)a_slabs * slab;
  const uint32_t stage_bytes = y_bytes + a_bytes;
  float* ones = reinterpret_cast<float*>(smem);  // 8 (16) k-rows x 128 B of 1.0 (B operand of the bias MMA)
  smem += 2048;
  uint8_t* ctl_raw = smem + (size_t)p.stages * stage_bytes;
  SmemCtlW* ctl = reinterpret_cast<SmemCtlW*>(ctl_raw);
  float* xpose = reinterpret_cast<float*>(ctl_raw + ((sizeof(SmemCtlW) + 15) & ~size_t(15)));
  const uint32_t crank = kMc ? cluster_ctarank() : 0u;
  const int n_items = kMc ? (p.n_tiles + 1) / 2 : p.n_tiles;  // n tiles (pairs of n tiles) per (k tile, split)
  const int total = n_items * p.k_tiles * p.splits;
  const int first = kMc ? (int)blockIdx.x / 2 : (int)blockIdx.x, step = kMc ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int h0 = (a_slabs + 1) / 2;  // kMc: A slabs loaded (and multicast) per CTA; an odd middle slab is delivered twice
  const uint32_t stage_tx = kMc ? y_bytes + 2u * (uint32_t)h0 * slab : stage_bytes;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmY);
    prefetch_tmap(&tmA);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&ctl->full[s]), 1);
      mbar_init(smem_u32(&ctl->empty[s]), kMc ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&ctl->tfull[i]), 1);
      mbar_init(smem_u32(&ctl->tempty[i]), kEpiWarps);
      mbar_init(smem_u32(&ctl->sfull[i]), 1);
      mbar_init(smem_u32(&ctl->sempty[i]), 1 + kEpiWarps);
    }
    fence_barrier_init();
  }
  const bool dyn = !kMc && p.work != nullptr;  // first item blockIdx.x, the following ones drawn (see gemm_tc_body)
  const int t_first = first;
  if (warp == 2) tmem_alloc(smem_u32(&ctl->tmem_base), kTmemCols);
  if (warp == 3) {
    if (kBf16) for (int i = lane; i < 512; i += 32) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;  // two bf16 1.0
    else for (int i = lane; i < 256; i += 32) ones[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
  }
  tc_fence_before();
  if (kMc) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  // item -> (split, k tile, n tile): consecutive CTAs share the row range (A / dY slabs hit in L2)
  auto decode = [&](int t, int& nt, int& kt, int& g0, int& g1) {
    const int tiles = n_items * p.k_tiles;
    const int sp = t / tiles, r = t - sp * tiles;
    kt = r / n_items;
    nt = r - kt * n_items;
    if (kMc) nt = 2 * nt + (int)crank;  // may be == n_tiles (odd count): dY slabs out of bounds -> zeros, nothing stored
    g0 = sp * p.gps;
    g1 = min(p.groups, g0 + p.gps);
  };

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      ItemWriter iw;
      iw.init(ctl);
      for (int t = t_first;;) {
        if (dyn) iw.put(t < total ? t : -1);
        if (t >= total) {
          if (dyn) work_leave(p.work);
          break;
        }
        const int t_next = dyn ? step + atomicAdd(p.work, 1) : t + step;
        int nt, kt, g0, g1;
        decode(t, nt, kt, g0, g1);
        tc_trace_w(p, 0, 0, t);
        for (int g = g0; g < g1; ++g) {
          const int bt_i = g / p.lt, lt_i = g - bt_i * p.lt;
          const int l0 = lt_i * p.bl, b0 = bt_i * p.nb;
          mbar_wait(smem_u32(&ctl->empty[s]), ph ^ 1);
          if (g == g0 || g == g1 - 1) tc_trace_w(p, 0, 1 + (g != g0), t);
          const uint32_t fb = smem_u32(&ctl->full[s]);
          mbar_expect_tx(fb, stage_tx);
          const uint32_t sy = smem_u32(smem + (size_t)s * stage_bytes);
          // one TMA per operand: 4-D boxes (32 floats, bl, nb, slabs) land slab-major, exactly the MN-major layout
          tma_load_4d(sy, &tmY, fb, 0, l0, b0, nt * sub * kYSlabs);
          if (kMc) {
            const int sl0 = crank ? a_slabs - h0 : 0;  // my half of the A slabs, into both CTAs' stage
            tma_load_4d_mc(sy + y_bytes + (uint32_t)sl0 * slab, &tmA, fb, 0, l0, b0, kt * (p.bnk / SW) + sl0, (uint16_t)3);
          } else
          tma_load_4d(sy + y_bytes, &tmA, fb, 0, l0, b0, kt * (p.bnk / SW));
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        t = t_next;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = kBf16 ? idesc_bf16(kBM, p.bnk, 1, 1) : idesc_tf32(kBM, p.bnk, 1, 1);
      const uint32_t idesc_b = kBf16 ? idesc_bf16(kBM, 16, 1, 1) : idesc_tf32(kBM, 16, 1, 1);
      const uint64_t ones_desc = kBf16 ? smem_desc(smem_u32(ones), 2048, 1024, 2) : smem_desc(smem_u32(ones), 1024, 512, 1);
      const uint32_t ones_lo = (uint32_t)ones_desc, ones_hi = (uint32_t)(ones_desc >> 32);
      // MN-major: LBO = slab pitch; fp32: 128B swizzle with 32-byte atoms, SBO = 512 B (4 k-rows); bf16: plain 128B
      // swizzle, SBO = 1024 B (8 k-rows)
      const uint64_t ydesc0 = kBf16 ? smem_desc(smem_u32(smem), slab, 1024, 2) : smem_desc(smem_u32(smem), slab, 512, 1);
      constexpr uint32_t kAdv = kRowsPerMma * 128 / 16;  // descriptor units per MMA (8 or 16 reduction rows of 128 B)
      const uint32_t ydesc_lo0 = (uint32_t)ydesc0, desc_hi = (uint32_t)(ydesc0 >> 32);
      const uint32_t stage_units = stage_bytes >> 4;
      int s = 0, it = 0;
      uint32_t ph = 0;
      ItemReader ir;
      ir.init(ctl);
      for (int t = first;; ++it) {
        if (dyn) t = ir.next(false, 0); else if (it) t += step;
        if (t < 0 || t >= total) break;
        int nt, kt, g0, g1;
        decode(t, nt, kt, g0, g1);
        const bool do_bias = p.dbias != nullptr && kt == 0;
        const int acc = sub == 2 ? 0 : (it & 1);
        tc_trace_w(p, 1, 0, t);
        mbar_wait(smem_u32(&ctl->tempty[acc]), (sub == 2 ? (it & 1) : ((it >> 1) & 1)) ^ 1);
        tc_fence_after();
        tc_trace_w(p, 1, 1, t);
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kMaxBN;
        for (int g = g0; g < g1; ++g) {
          mbar_wait(smem_u32(&ctl->full[s]), ph);
          tc_fence_after();
          if (g == g0 || g == g1 - 1) tc_trace_w(p, 1, 2 + (g != g0), t);
          const uint32_t sy = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sa = sy + y_bytes;
          const uint32_t y_lo = ydesc_lo0 + (uint32_t)s * stage_units, a_lo = y_lo + (y_bytes >> 4);
          uint32_t accf = g > g0 ? 1u : 0u;
          for (int r8 = 0; r8 < R / kRowsPerMma; ++r8) {  // 8 (16) reduction rows = 1024 (2048) B per MMA
            umma_lh<kBf16>(d_tmem, y_lo + r8 * kAdv, desc_hi, a_lo + r8 * kAdv, desc_hi, idesc, accf);
            if (do_bias) umma_lh<kBf16>(d_tmem + kBiasCol, y_lo + r8 * kAdv, desc_hi, ones_lo, ones_hi, idesc_b, accf);
            if (sub == 2) {  // second n tile: the next dY slabs of the same stage, same A slabs
              const uint32_t y2 = y_lo + ((kYSlabs * slab) >> 4) + r8 * kAdv;
              umma_lh<kBf16>(d_tmem + kMaxBN, y2, desc_hi, a_lo + r8 * kAdv, desc_hi, idesc, accf);
              if (do_bias) umma_lh<kBf16>(d_tmem + kMaxBN + kBiasCol, y2, desc_hi, ones_lo, ones_hi, idesc_b, accf);
            }
            accf = 1u;
          }
          if (kMc) umma_commit_mc(smem_u32(&ctl->empty[s]), (uint16_t)3); else umma_commit(smem_u32(&ctl->empty[s]));
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        umma_commit(smem_u32(&ctl->tfull[acc]));
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int ew = warp - 4, q = ew & 3, h = ew >> 2;
    constexpr int kColStep = 32 * (kEpiWarps / 4);
    const int rq = lane >> 3, cq = lane & 7;
    float4* xp4 = reinterpret_cast<float4*>(xpose + ew * kXposeFloats);
    int it = 0;
    ItemReader ir;
    ir.init(ctl);
    for (int t = first;; ++it) {
      if (dyn) t = ir.next(true, lane); else if (it) t += step;
      if (t < 0 || t >= total) break;
      int nt, kt, g0, g1;
      decode(t, nt, kt, g0, g1);
      const int acc = sub == 2 ? 0 : (it & 1);
      if (ew == 0 && lane == 0) tc_trace_w(p, 2, 0, t);
      mbar_wait(smem_u32(&ctl->tfull[acc]), sub == 2 ? (it & 1) : ((it >> 1) & 1));
      tc_fence_after();
      if (ew == 0 && lane == 0) tc_trace_w(p, 2, 1, t);
      const int kbase = kt * p.bnk;
      const int kcols = min(p.bnk, p.K - kbase);
     for (int sj = 0; sj < sub; ++sj) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub == 2 ? sj : acc) * kMaxBN;
      const int nbase = (nt * sub + sj) * kBM + q * 32;
      if (g1 > g0 && nbase < p.N) {
        const int nrows = min(32, p.N - nbase);
        for (int c0 = h * 32; c0 < kcols; c0 += kColStep) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            xp4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          __syncwarp();
          const int cc = c0 + cq * 4;
          if (cc < kcols) {  // K % 4 == 0: whole float4s
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + rq;
              if (r < nrows) red_add_v4(p.dW + (size_t)(nbase + r) * p.K + kbase + cc, xp4[r * 8 + (cq ^ (r & 7))]);
            }
          }
          __syncwarp();
        }
        if (p.dbias != nullptr && kt == 0 && h == 0) {  // every one of the 16 bias columns holds sum_rows dY[row][n] for n = TMEM lane
          uint32_t bv[16];
          tmem_ld16(taddr + kBiasCol, bv);
          tmem_ld_wait();
          const int n = nbase + lane;
          if (n < p.N && n < p.bias_n) atomicAdd(p.dbias + (n % p.bias_mod), __uint_as_float(bv[0]));
        }
      }
     }  // n tiles of the item
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ctl->tempty[acc]));
      if (ew == 0 && lane == 0) tc_trace_w(p, 2, 2, t);
    }
  }
  tc_fence_before();
  if (kMc) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmA, const WgradTcParams p) {
  wgrad_tc_body<false>(tmY, tmA, p);
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_bf16_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmA, const WgradTcParams p) {
  wgrad_tc_body<false, true>(tmY, tmA, p);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
wgrad_tc_mc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmA, const WgradTcParams p) {
  wgrad_tc_body<true>(tmY, tmA, p);
}

// counters of the dynamically scheduled launches: a ring of (next item, CTAs done) pairs, zero at load and reset by the
// last CTA of every launch; consecutive launches take consecutive slots, so kernels running concurrently on different
// streams (or as parallel branches of one CUDA graph) never share one.  A slot is baked into a captured graph node; the
// ring (4096 slots, ~25 training steps' worth of launches) would only hand the same slot to two kernels that RUN AT THE
// SAME TIME if a graph replay overlapped eager launches made 4096 x k launches after its capture on another stream —
// the engine serialises a step's graph and its eager work on one stream order, so that does not occur
constexpr int kWorkSlots = 4096;
__device__ int g_work_slots[2 * kWorkSlots];
int* g_work_base[64] = {};

int* work_slot() {
  const char* e = getenv("SCV_TC_DYN");  // read per call: SCV_TC_DYN=0 selects the static round robin
  if (e && atoi(e) == 0) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_work_base[dev]) return nullptr;
  static std::atomic<unsigned> seq{0};
  return g_work_base[dev] + 2 * (seq.fetch_add(1u) % (unsigned)kWorkSlots);
}

int g_attr_done = 0;
int g_pair_capacity = -1;
int g_mc_capacity = -1;

// why the tensor-core path declined a shape (falls back to the fp32 FFMA kernels): logged once per distinct reason
int decline(const char* who, const char* why, int64_t M, int64_t N, int64_t K) {
  static int logged = 0;
  static const bool quiet = [] { const char* e = getenv("SCV_QUIET"); return e && atoi(e); }();
  if (!quiet && M * N * K >= (1LL << 24) && logged < 8) {  // tiny problems (scrubber heads) are FFMA by design
    ++logged;
    fprintf(stderr, "libscv: %s takes the fp32 FFMA path for M=%lld N=%lld K=%lld (%s)\n", who, (long long)M, (long long)N,
            (long long)K, why);
  }
  return 1;
}

// number of CTA pairs (clusters of 2) of the 2-CTA GEMM that can be resident at once
int pair_capacity() { return g_pair_capacity; }

int ensure_attrs() {
  if (g_attr_done) return 0;
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_bnr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_bnr_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (scv::sm_count() / 2));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemLimit;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tc2_kernel, &cfg) == cudaSuccess) g_pair_capacity = n;
    else { g_pair_capacity = 0; cudaGetLastError(); }
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_mc_kernel, &cfg) == cudaSuccess) g_mc_capacity = n;
    else { g_mc_capacity = 0; cudaGetLastError(); }
  }
  if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e != cudaSuccess) {
    scv::set_error("tensor-core GEMM: cannot opt in to %d B of shared memory: %s", kSmemLimit, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  {
    int dev = 0;
    void* sym = nullptr;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && cudaGetSymbolAddress(&sym, g_work_slots) == cudaSuccess)
      g_work_base[dev] = static_cast<int*>(sym);
    else
      cudaGetLastError();  // no dynamic scheduling on this device: the static round robin still works
  }
  g_attr_done = 1;
  return 0;
}

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// rows of a (bl x nb) box, at most `rmax`; forward wants the best use of the 128 MMA rows,
// the weight gradient wants rows % 8 == 0 and as few zero-filled rows as possible
void choose_box(int64_t Lo, int64_t B, int rmax, int mult, int& bl, int& nb) {  // mult: rows must be a multiple of it (0: any)
  const bool mult8 = mult > 1;
  double best = -1.0;
  bl = 1;
  nb = mult8 ? mult : 1;
  int lmax = (int)(mult8 ? (Lo + mult - 1) / mult * mult : Lo);
  if (lmax > rmax) lmax = rmax;
  if (lmax > 256) lmax = 256;
  for (int l = lmax; l >= 1; --l) {
    for (int n = 1; n * l <= rmax && n <= 256; ++n) {
      const int r = l * n;
      if (mult8 && r % mult) continue;
      const double covered = (double)(cdiv(Lo, l) * l) * (double)(cdiv(B, n) * n);
      double eff = (double)Lo * (double)B / covered;
      eff *= (double)r / rmax;  // forward: unused MMA rows; wgrad: short stages pay more barrier round trips
      if (eff > best + 1e-9 || (eff > best - 1e-9 && r > bl * nb)) {
        best = eff;
        bl = l;
        nb = n;
      }
    }
  }
}

}  // namespace

namespace scv {

int gemm_tc(const scv_gemm_t* p, cudaStream_t st) {
  const int64_t M = p->B * p->Lo;
  const bool bf16 = p->precision == SCV_PREC_BF16;
  const int ea = bf16 ? 8 : 4;  // operand elements per 16 bytes
  // shapes the tensor-core path does not take (tiny scrubber-head layers, unaligned views); bf16 operands have no FFMA
  // fallback, so only the hard limits apply to them
  if (p->N < 16 || p->K < (bf16 ? 8 : 32) || (!bf16 && M < 64))
    return decline("scv_gemm", "N < 16, K < 32 or fewer than 64 rows", M, p->N, p->K);
  if (p->K % ea || p->a_bs % ea || p->a_ls % ea || !aligned16(p->A) || !aligned16(p->W))
    return decline("scv_gemm", "A / W not 16-byte aligned", M, p->N, p->K);
  if (p->K > (1 << 30) || p->N > (1 << 30) || p->B >= (1LL << 31) || p->Lo >= (1LL << 31))
    return decline("scv_gemm", "extent beyond 2^30", M, p->N, p->K);
  // the epilogue moves float4s: rows of Y / R / bias must be 16-byte aligned at every 4-column group
  if (p->N % 4 || p->n_last % 4 || p->y_bs % 4 || p->y_ls % 4 || !aligned16(p->Y))
    return decline("scv_gemm", "Y rows not float4-aligned", M, p->N, p->K);
  if (p->R && (p->r_bs % 4 || p->r_ls % 4 || !aligned16(p->R))) return decline("scv_gemm", "R rows not float4-aligned", M, p->N, p->K);
  if (p->bias && (p->bias_mod % 4 || p->bias_n % 4 || !aligned16(p->bias)))
    return decline("scv_gemm", "bias not float4-aligned", M, p->N, p->K);
  const bool bnr = p->bnr_sums != nullptr;
  if (bnr && ((p->act & 15) != SCV_ACT_NONE || (p->act & (SCV_ACT_ROUND_TF32 | SCV_ACT_ACCUM)) || p->stats || !p->bnr_x ||
              p->bnr_c <= 0 || p->bnr_c % 4 || p->N % p->bnr_c || p->bnr_bs != p->y_bs || (p->Lo > 1 && p->bnr_ls != p->y_ls) ||
              (p->R && (p->r_bs != p->y_bs || (p->Lo > 1 && p->r_ls != p->y_ls))) || !aligned16(p->bnr_x) ||
              (p->bnr_chan && !aligned16(p->bnr_chan))))
    return decline("scv_gemm", "fused BatchNorm backward reduction needs a plain epilogue and X (and R) laid out like Y", M, p->N, p->K);
  const bool accum = (p->act & SCV_ACT_ACCUM) != 0;
  if (accum && ((p->act & 15) != SCV_ACT_NONE || (p->act & SCV_ACT_ROUND_TF32) || p->R || p->stats))
    return decline("scv_gemm", "SCV_ACT_ACCUM with an activation / residual / statistics", M, p->N, p->K);
  int rc = ensure_attrs();
  if (rc) return rc;

  GemmTcParams q;
  q.B = p->B; q.Lo = p->Lo; q.K = (int)p->K; q.N = (int)p->N;
  choose_box(p->Lo, p->B, kBM, 0, q.bl, q.nb);
  q.lt = (int)cdiv(p->Lo, q.bl);
  q.bt = (int)cdiv(p->B, q.nb);
  q.m_tiles = q.lt * q.bt;
  q.k_chunks = (int)cdiv(p->K, bf16 ? 64 : kBK);
  // CTA pairs (cta_group::2, gemm_tc2_kernel).  Round 1 measured them slower on every layer (a pipeline bug: the peer CTA
  // arrived on the leader's barrier every chunk) and left them off.
  // Round 2: with the issue paths on elect.sync (and the peer's per-chunk remote arrive gone) the pair kernel wins on the
  // large layers — a CTA pair ingests a third less operand data per FLOP.  Measured per layer (gpurun_out/gemm_r2q*.json,
  // profiles/r02_pair_vs_single.md): N >= 128 and K >= 1024 picks every layer with a gain (-115 of 2524 us per step over
  // the forward + data-gradient GEMMs) and none that loses more than 1 us.  SCV_TC_PAIR=0 / 1 forces never / always.
  const char* pair_env = getenv("SCV_TC_PAIR");  // read per call (tests switch it)
  const int force_pair = pair_env ? atoi(pair_env) : -1;
  int ctas = force_pair == 1 ? 2 : (force_pair == 0 ? 1 : ((p->N >= 128 && p->K >= 1024) ? 2 : 1));
  if (q.m_tiles < 2 || pair_capacity() < 1 || bf16 || bnr || accum) ctas = 1;
  // N tile: as wide as the MMA allows, balanced over the tiles, multiple of 16 (of 32 for a pair: each CTA holds half)
  const int ng = 16 * ctas;
  const int n16 = (int)cdiv(p->N, ng) * ng;
  const int nt0 = (int)cdiv(n16, kMaxBN);
  q.bn = (int)cdiv(cdiv(n16, nt0), ng) * ng;
  q.n_tiles = (int)cdiv(p->N, q.bn);
  // two row tiles per work item when that still fills the machine and K is long enough for the shared-memory
  // traffic (not the epilogue, which is no longer overlapped) to dominate; SCV_TC_SUB=1/2 forces the choice
  const int force_sub = [] { const char* e = getenv("SCV_TC_SUB"); return e ? atoi(e) : 0; }();  // per call (tests)
  q.sub = 1;
  if (ctas == 1) {
    const int items2 = q.n_tiles * (int)cdiv(q.m_tiles, 2);
    // measured rule (profiles/r01_sub2_vs_sub1.md): enough items for 3/4 of the SMs, and a main loop (8 MMAs per
    // 32-float K chunk, >= 48 cycles each) at least 1.5x the two un-overlapped epilogues (~1500 cycles per 32 columns)
    const int64_t mainloop = (int64_t)q.k_chunks * 8 * (q.bn / 2 > 48 ? q.bn / 2 : 48);
    const int64_t epi = 2 * cdiv(q.bn, 32) * 1500;
    if (items2 >= (sm_count() * 3) / 4 && 2 * mainloop >= 3 * epi) q.sub = 2;
    if (force_sub == 1 || force_sub == 2) q.sub = force_sub;
    if (q.m_tiles < 2) q.sub = 1;
  }
  const size_t fixed = 1024 + ((sizeof(SmemCtl) + 15) & ~size_t(15)) + kEpiWarps * kXposeFloats * 4 +
                       ((p->stats || p->bnr_sums) ? 2 * kMaxBN * sizeof(float) : 0);
  size_t stage_bytes = (size_t)kBM * kBK * 4 * q.sub + (size_t)(q.bn / ctas) * kBK * 4;
  if (q.sub == 2 && (kSmemLimit - fixed) / stage_bytes < 3 && force_sub != 2) {
    q.sub = 1;  // two stages cannot hide the TMA latency (measured: 256 x 256 items with 2 stages lose 30 %)
    stage_bytes = (size_t)kBM * kBK * 4 + (size_t)(q.bn / ctas) * kBK * 4;
  }
  // W multicast across a cluster of two CTAs (same n tile, adjacent row-tile groups).  MEASURED: no gain — 2-4 % slower on
  // every layer (profiles/r02_multicast_vs_unicast.md): each SM still has to take delivery of the full W tile, and L2
  // already merges the two CTAs' unicast requests.  Off by default; SCV_TC_MC=1 selects it (the kernel tests cover both).
  const int mc_env = [] { const char* e = getenv("SCV_TC_MC"); return e ? atoi(e) : 0; }();  // read per call: tests toggle it
  const bool mc = ctas == 1 && !bf16 && !bnr && mc_env != 0 && g_mc_capacity >= 1 && q.m_tiles >= 2 * q.sub && q.bn % 16 == 0;
  // split-K (accumulating GEMMs only): few output tiles, long reduction
  q.ksplit = 1;
  q.kc_per = q.k_chunks;
  if (accum && ctas == 1) {
    const int tiles0 = q.n_tiles * (int)cdiv(q.m_tiles, q.sub * (mc ? 2 : 1)) * (mc ? 2 : 1);
    int want = sm_count() / (tiles0 > 0 ? tiles0 : 1);
    if (want > q.k_chunks / 4) want = q.k_chunks / 4;
    if (want > 1) {
      q.kc_per = (int)cdiv(q.k_chunks, want);
      q.ksplit = (int)cdiv(q.k_chunks, q.kc_per);
    }
  }
  int stages = (int)((kSmemLimit - fixed) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > q.kc_per + 1) stages = q.kc_per + 1;
  if (stages < 2) stages = 2;
  q.stages = stages;
  q.bias = p->bias; q.bias_mod = (int)(p->bias ? p->bias_mod : 1); q.bias_n = (int)(p->bias ? p->bias_n : 0);
  q.Y = p->Y; q.y_bs = p->y_bs; q.y_ls = p->y_ls; q.n_last = (int)p->n_last;
  q.R = p->R; q.r_bs = p->r_bs; q.r_ls = p->r_ls;
  q.act = (int)(p->act & 15); q.round_out = (p->act & SCV_ACT_ROUND_TF32) ? 1 : 0; q.out_scale = (float)p->out_scale; q.stats = p->stats;
  {
    const bool rd = q.R != nullptr, stt = q.stats != nullptr, rn = q.round_out != 0;
    if (bnr) q.epi_kind = rd ? 8 : 7;
    else if (accum) q.epi_kind = 6;
    else if (q.act == SCV_ACT_NONE && !rn) q.epi_kind = (rd ? 2 : 0) + (stt ? 1 : 0);
    else if (q.act == SCV_ACT_NONE && rn && !rd && !stt) q.epi_kind = 4;
    else if (q.act == SCV_ACT_TANH && !rn && !rd && !stt) q.epi_kind = 5;
    else q.epi_kind = 99;  // generic
  }
  q.bnr_x = p->bnr_x; q.bnr_bs = p->bnr_bs; q.bnr_ls = p->bnr_ls; q.bnr_chan = p->bnr_chan; q.bnr_slope = p->bnr_slope;
  q.bnr_c = (int)p->bnr_c; q.bnr_sums = p->bnr_sums;
  q.work = nullptr;
  q.trace = nullptr;
  if (const char* tr = getenv("SCV_TC_TRACE")) {  // debug: device buffer address (decimal) to receive CTA 0's timeline
    q.trace = reinterpret_cast<long long*>(strtoull(tr, nullptr, 10));
  }

  CUtensorMap tmA, tmW;
  {
    const int64_t dims[3] = {p->K, p->Lo, p->B};
    const int64_t str[3] = {1, p->a_ls ? p->a_ls : p->a_bs, p->a_bs ? p->a_bs : ea};
    const int box[3] = {bf16 ? 64 : kBK, q.bl, q.nb};
    rc = tc::make_tmap(&tmA, p->A, 3, dims, str, box, "scv_gemm A", false, bf16);
    if (rc) return rc;
  }
  {
    const int64_t dims[2] = {p->K, p->N};
    const int64_t str[2] = {1, p->K};
    const int box[2] = {bf16 ? 64 : kBK, q.bn / (mc ? 2 : ctas)};
    rc = tc::make_tmap(&tmW, p->W, 2, dims, str, box, "scv_gemm W", false, bf16);
    if (rc) return rc;
  }
  const size_t smem = fixed + (size_t)stages * stage_bytes;
  if (ctas == 2) {
    const int items = q.n_tiles * (int)cdiv(q.m_tiles, 2);
    const int pairs = items < pair_capacity() ? items : pair_capacity();
    gemm_tc2_kernel<<<2 * pairs, kThreads, smem, st>>>(tmA, tmW, q);
    return check_launch("gemm_tc2_kernel");
  }
  if (mc) {
    const int items = q.n_tiles * (int)cdiv(q.m_tiles, 2 * q.sub) * q.ksplit;
    const int pairs = items < g_mc_capacity ? items : g_mc_capacity;
    gemm_tc_mc_kernel<<<2 * pairs, kThreads, smem, st>>>(tmA, tmW, q);
    return check_launch("gemm_tc_mc_kernel");
  }
  const int total = q.n_tiles * (int)cdiv(q.m_tiles, q.sub) * q.ksplit;
  const int grid = total < sm_count() ? total : sm_count();
  q.work = total > grid ? work_slot() : nullptr;
  if (bnr) {
    if (bf16) gemm_tc_bnr_bf16_kernel<<<grid, kThreads, smem, st>>>(tmA, tmW, q);
    else gemm_tc_bnr_kernel<<<grid, kThreads, smem, st>>>(tmA, tmW, q);
    return check_launch("gemm_tc_bnr_kernel");
  }
  if (bf16) {
    gemm_tc_bf16_kernel<<<grid, kThreads, smem, st>>>(tmA, tmW, q);
    return check_launch("gemm_tc_bf16_kernel");
  }
  gemm_tc_kernel<<<grid, kThreads, smem, st>>>(tmA, tmW, q);
  return check_launch("gemm_tc_kernel");
}

int wgrad_tc(const scv_wgrad_t* p, cudaStream_t st) {
  const int64_t M = p->B * p->Lo;
  const bool bf16 = p->precision == SCV_PREC_BF16;
  const int ea = bf16 ? 8 : 4, SW = bf16 ? 64 : 32;
  if (p->N < 16 || p->K < (bf16 ? 8 : 32) || (!bf16 && M < 256))
    return decline("scv_wgrad", "N < 16, K < 32 or fewer than 256 rows", M, p->N, p->K);
  if (p->K % 4 || p->N % 4 || p->a_bs % ea || p->a_ls % ea || p->y_bs % ea || p->y_ls % ea || !aligned16(p->A) ||
      !aligned16(p->dY) || !aligned16(p->dW))
    return decline("scv_wgrad", "operands not 16-byte aligned", M, p->N, p->K);
  if (p->K > (1 << 30) || p->N > (1 << 30) || p->B >= (1LL << 31) || p->Lo >= (1LL << 31))
    return decline("scv_wgrad", "extent beyond 2^30", M, p->N, p->K);
  int rc = ensure_attrs();
  if (rc) return rc;

  WgradTcParams q;
  q.K = (int)p->K; q.N = (int)p->N; q.dW = p->dW;
  const bool want_bias = p->dbias && p->bias_n > 0;
  q.dbias = want_bias ? p->dbias : nullptr;
  q.bias_mod = (int)(want_bias ? p->bias_mod : 1);
  q.bias_n = (int)(want_bias ? p->bias_n : 0);
  q.work = nullptr;
  q.trace = nullptr;
  if (const char* tr = getenv("SCV_TC_TRACE")) q.trace = reinterpret_cast<long long*>(strtoull(tr, nullptr, 10));
  const int wrows_env = [] { const char* e = getenv("SCV_TC_WROWS"); return e ? atoi(e) : 0; }();  // experiments
  choose_box(p->Lo, p->B, wrows_env >= 8 ? wrows_env : (bf16 ? 64 : 32), bf16 ? 16 : 8, q.bl, q.nb);
  q.lt = (int)cdiv(p->Lo, q.bl);
  q.bt = (int)cdiv(p->B, q.nb);
  q.groups = q.lt * q.bt;
  // k tile: whole 32-float slabs (the TMA box counts slabs), balanced over the tiles
  const int k32 = (int)cdiv(p->K, SW) * SW;
  const int wbnk_env = [] { const char* e = getenv("SCV_TC_WBNK"); return e ? atoi(e) : 0; }();  // experiments
  const int kt0 = (int)cdiv(k32, wbnk_env >= 32 ? wbnk_env : (bf16 ? 192 : kWgradMaxBNK));
  q.bnk = (int)cdiv(cdiv(k32, kt0), SW) * SW;
  q.k_tiles = (int)cdiv(p->K, q.bnk);
  // two n tiles per item (sharing the A slabs) measured 7-10 % SLOWER than one on every large layer
  // (profiles/r01_wgrad_sub.md: fewer stages in flight and an un-overlapped epilogue cost more than the saved
  // traffic), so one tile per item is the default; SCV_TC_WSUB=2 selects the two-tile variant for experiments
  static const int force_wsub = [] { const char* e = getenv("SCV_TC_WSUB"); return e ? atoi(e) : 0; }();
  q.sub = (force_wsub == 2 && !bf16) ? 2 : 1;
  if (p->N <= kBM) q.sub = 1;
  q.n_tiles = (int)cdiv(p->N, kBM * q.sub);
  // A multicast across a cluster of two CTAs holding adjacent n tiles: measured 0-4 % slower (same profile); SCV_TC_WMC=1
  const int wmc_env = [] { const char* e = getenv("SCV_TC_WMC"); return e ? atoi(e) : 0; }();
  const bool mc = wmc_env != 0 && !bf16 && g_mc_capacity >= 1 && q.sub == 1 && q.n_tiles >= 2 && q.bnk >= 64;
  const int tiles = (mc ? (q.n_tiles + 1) / 2 * 2 : q.n_tiles) * q.k_tiles;
  // row splits: fill the SMs ~4x over (measured on the 28 weight-gradient GEMMs of the default network, B = 2048:
  // 1x 1800 us, 2x 1553 us, 3x 1554 us, 4x 1487 us; shorter items even out the tail), once over for the two-tile
  // items (their epilogue is not overlapped, so fewer, longer items), but keep at least 4 row groups per item
  const int wsplit_env = [] { const char* e = getenv("SCV_TC_WSPLIT"); return e ? atoi(e) : 0; }();  // experiments
  int splits = (int)cdiv((wsplit_env > 0 ? wsplit_env : (q.sub == 2 ? 1 : 4)) * sm_count(), tiles);
  if (splits > q.groups / 4) splits = q.groups / 4;
  if (splits < 1) splits = 1;
  q.gps = (int)cdiv(q.groups, splits);
  q.splits = (int)cdiv(q.groups, q.gps);
  const int R = q.bl * q.nb;
  const size_t stage_bytes = (size_t)(kBM / SW * q.sub + (q.bnk + SW - 1) / SW) * R * 128;
  const size_t fixed = 1024 + 2048 + ((sizeof(SmemCtlW) + 15) & ~size_t(15)) + kEpiWarps * kXposeFloats * 4;
  int stages = (int)((kSmemLimit - fixed) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  const int wstages_env = [] { const char* e = getenv("SCV_TC_WSTAGES"); return e ? atoi(e) : 0; }();  // experiments
  if (wstages_env >= 2 && stages > wstages_env) stages = wstages_env;
  if (stages < 2) return decline("scv_wgrad", "a stage does not fit shared memory twice", M, p->N, p->K);
  q.stages = stages;

  CUtensorMap tmY, tmA;
  // (32 floats, l, b, slab): the 4th dimension walks the 32-float column slabs of a row.  A partial last slab
  // reads past the row's end (the next row / the buffer's slack): those columns only reach output rows n >= N or
  // output columns k >= K, which the epilogue does not store.
  {
    const int64_t dims[4] = {SW, p->Lo, p->B, cdiv(p->N, SW)};
    const int64_t str[4] = {1, p->y_ls ? p->y_ls : p->y_bs, p->y_bs ? p->y_bs : ea, SW};
    const int box[4] = {SW, q.bl, q.nb, kBM / SW * q.sub};
    rc = tc::make_tmap(&tmY, p->dY, 4, dims, str, box, "scv_wgrad dY", !bf16, bf16);
    if (rc) return rc;
  }
  {
    const int64_t dims[4] = {SW, p->Lo, p->B, cdiv(p->K, SW)};
    const int64_t str[4] = {1, p->a_ls ? p->a_ls : p->a_bs, p->a_bs ? p->a_bs : ea, SW};
    const int box[4] = {SW, q.bl, q.nb, mc ? (q.bnk / SW + 1) / 2 : q.bnk / SW};
    rc = tc::make_tmap(&tmA, p->A, 4, dims, str, box, "scv_wgrad A", !bf16, bf16);
    if (rc) return rc;
  }
  const size_t smem = fixed + (size_t)stages * stage_bytes;
  if (mc) {
    const int items = (q.n_tiles + 1) / 2 * q.k_tiles * q.splits;
    const int pairs = items < g_mc_capacity ? items : g_mc_capacity;
    wgrad_tc_mc_kernel<<<2 * pairs, kThreads, smem, st>>>(tmY, tmA, q);
    return check_launch("wgrad_tc_mc_kernel");
  }
  const int total = tiles * q.splits;
  const int grid = total < sm_count() ? total : sm_count();
  q.work = total > grid ? work_slot() : nullptr;
  if (bf16) {
    wgrad_tc_bf16_kernel<<<grid, kThreads, smem, st>>>(tmY, tmA, q);
    return check_launch("wgrad_tc_bf16_kernel");
  }
  wgrad_tc_kernel<<<grid, kThreads, smem, st>>>(tmY, tmA, q);
  rc = check_launch("wgrad_tc_kernel");
  if (rc) return rc;
  return rc;
}

}  // namespace scv
