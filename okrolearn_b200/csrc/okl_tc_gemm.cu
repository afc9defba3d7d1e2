// tcgen05 tensor-core GEMM for sm_100a: D[M,N] = act(A . B + bias), TF32 or BF16 operands, FP32 accumulation in TMEM.
//
// Used for DenseLayer forward / dgrad / wgrad (okl.py:29-47) in the OKL_TF32 / OKL_BF16 modes.  All three products run
// on the tensors exactly as the reference lays them out (X (M,K), W (K,N), G (M,N), all row-major) - the operand
// "major-ness" is expressed in the UMMA descriptors instead of transposing anything:
//     fwd   Y  = X  . W      A = X  K-major,   B = W  MN-major
//     dgrad dX = G  . W^T    A = G  K-major,   B = W  K-major
//     wgrad dW = X^T . G     A = X  MN-major,  B = G  MN-major
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//     warp 0    TMA producer : cp.async.bulk.tensor (128B-swizzled boxes) into a multi-stage smem ring, mbarrier expect_tx
//     warp 1    MMA issuer   : one elected thread issues tcgen05.mma (M=128, N=BN, K=8 tf32 / 16 bf16) into TMEM,
//                              tcgen05.commit releases smem stages and publishes the accumulator
//     warp 2    TMEM allocator (2 x BN columns: the epilogue of tile i overlaps the main loop of tile i+1)
//     warps 4-7 epilogue     : tcgen05.ld -> registers -> bias + activation -> 128-bit global stores
// Small problems are split along K over CTAs (fp32 partials in a workspace, reduced in fixed order by a second kernel
// that applies the epilogue) so that e.g. the 256x5408x128 DenseLayer of the MNIST CNN still fills the 148 SMs.
#include "common.cuh"
#include "conv_geom.cuh"

#include <cuda_bf16.h>

namespace {

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 384;        // 4 control warps (TMA, MMA, TMEM alloc, spare) + 8 epilogue warps

}  // namespace
#include "tc_ptx.cuh"
namespace {

struct TcGemmParams {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_per_split, kb_total;
  float* D;            // output (ldd) when splits == 1
  long long ldd;
  const float* bias;   // per-column bias or nullptr
  const float* mask;   // optional, same indexing as D: D = 0 where mask <= 0 (fused backward of a preceding ReLU)
  int act;
  float* ws;           // [splits][M][N] partials when splits > 1
  int stages;          // smem ring depth actually used (<= TcCfg::STAGES): short K loops take less smem so CTAs of concurrent kernels co-reside
  int* counters;       // [tiles] arrival counters (zero on entry, reset by the last arriver) when splits > 1
  // implicit-GEMM convolution (MODE 1 = fprop/dgrad, MODE 2 = wgrad); activations are channels-last (N, H, W, C)
  int cv_TW, cv_TH, cv_TN;         // spatial tile: TW x TH pixels of TN images (128 pixels = an M tile in MODE 1, 32 = a k-block in MODE 2)
  int cv_tw_sh, cv_th_sh;          // log2 of TW, TH
  int cv_tiles_w, cv_tiles_h;      // tiles per image
  int cv_OH, cv_OW, cv_NB;         // extents of the tiled pixel space and batch
  int cv_C, cv_cblocks, cv_kw, cv_ph, cv_pw;
  unsigned long long* dbg;   // optional per-CTA phase timestamps (globaltimer ns): start, setup, first full, acc ready, done
  __nv_bfloat16* D16;  // optional bf16 shadow of D (same ld), or nullptr
};

template <int KIND, int BN>
struct TcCfg {
  static constexpr int ELEM = KIND == 0 ? 4 : 2;
  static constexpr int BK = 128 / ELEM;                 // elements of K per stage (one 128-byte swizzle row)
  static constexpr int UMMA_K = 32 / ELEM;              // 8 (tf32) / 16 (bf16)
  static constexpr int EPC = 128 / ELEM;                // elements per 128-byte chunk along MN (MN-major operands)
  static constexpr int A_BYTES = TC_BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 8 * 16 * 36 * 4 /*epilogue staging*/;
};

// ------------------------------------------------------------------------------------------------ kernel
// PAIR = 1 (bf16, MODE 0, no split-K): CTA pairs with tcgen05 cta_group::2 - one 256 x BN MMA per pair, every CTA holds its 128 rows of
// A and HALF of the B tile, so the per-SM operand fetch from shared memory drops from (128 + BN) to (128 + BN / 2) rows per MMA.
template <int KIND, bool A_MN, bool B_MN, int BN, int MODE = 0, int PAIR = 0>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcGemmParams p) {
  using Cfg = TcCfg<KIND, BN>;
  static_assert(!PAIR || (KIND == 1 && MODE == 0), "pair mode: bf16 dense GEMM only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);
  volatile uint32_t* issue_par = reinterpret_cast<volatile uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 192);   // loop parameters of the MMA issuer
  float* stage_base = reinterpret_cast<float*>(smem + p.stages * Cfg::STAGE_BYTES + 256);   // 4 warps x 32 rows x 36 floats

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 0] = gtime();

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < p.stages; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PAIR ? 16 : 8);      // one arrive per epilogue warp (pair mode: of both CTAs, on the even CTA)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_cg2<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  if (threadIdx.x == 96) {
    issue_par[0] = (uint32_t)p.stages;
    issue_par[1] = p.dbg ? 1u : 0u;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  okl_pdl_sync();                          // barriers, TMEM and descriptor prefetch above overlap the predecessor's tail
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 1] = gtime();
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  if (PAIR) cluster_sync_all();
  constexpr int CS = PAIR ? 2 : 1;
  const int wstart = blockIdx.x / CS, wstep = gridDim.x / CS;       // pair mode: a work item is a PAIR of M tiles

  const int total_work = (p.tiles_m / CS) * p.tiles_n * p.splits;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = wstart; w < total_work; w += wstep) {
        const int tmg = p.tiles_m / CS;
        const int split = w / (tmg * p.tiles_n);
        const int t = w - split * (tmg * p.tiles_n);
        const int tm = (t % tmg) * CS + (int)crank, tn = t / tmg;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int m0 = tm * TC_BM, n0 = tn * BN;
        // implicit-GEMM wgrad: the k-block index walks (w tile, h tile, image group); decoded once per work item and then advanced
        // incrementally - the three divisions per k-block this used to cost (~400 dependent cycles in the single producer thread)
        // were as long as the 4 MMAs of a k-block
        int wg_wt = 0, wg_ht = 0, wg_ng = 0, wg_tap = 0, wg_c0 = 0, wg_u = 0, wg_v = 0;
        int f_tap = 0, f_cb = 0, f_u = 0, f_v = 0, f_wt = 0, f_ht = 0, f_ng = 0;
        if (MODE == 2) {
          wg_wt = kb0 % p.cv_tiles_w;
          const int r2 = kb0 / p.cv_tiles_w;
          wg_ht = r2 % p.cv_tiles_h; wg_ng = r2 / p.cv_tiles_h;
          wg_tap = n0 / p.cv_C; wg_c0 = n0 - wg_tap * p.cv_C;
          wg_u = wg_tap / p.cv_kw; wg_v = wg_tap - wg_u * p.cv_kw;
        }
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (!PAIR) mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          else if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::A_BYTES + Cfg::B_BYTES);
          const int k0 = kb * Cfg::BK;
          if (MODE == 1) {
            // implicit-GEMM fprop / dgrad: k-block = (filter tap, 32-channel chunk); the A tile is the TW x TH x TN pixel rectangle of
            // this M tile shifted by the tap - TMA zero-fills what falls outside the image, which IS the zero padding.
            if (kb == kb0) {
              f_tap = kb / p.cv_cblocks; f_cb = kb - f_tap * p.cv_cblocks;
              f_u = f_tap / p.cv_kw; f_v = f_tap - f_u * p.cv_kw;
              f_wt = tm % p.cv_tiles_w;
              const int r2 = tm / p.cv_tiles_w;
              f_ht = r2 % p.cv_tiles_h; f_ng = r2 / p.cv_tiles_h;
            }
            tma_load_4d(&tmA, &full_bar[stage], sa, f_cb * Cfg::BK, f_wt * p.cv_TW + f_v - p.cv_pw, f_ht * p.cv_TH + f_u - p.cv_ph, f_ng * p.cv_TN);
            tma_load_2d(&tmB, &full_bar[stage], sb, f_tap * p.cv_C + f_cb * Cfg::BK, n0);
            if (++f_cb == p.cv_cblocks) { f_cb = 0; f_tap++; if (++f_v == p.cv_kw) { f_v = 0; f_u++; } }
          } else if (MODE == 2) {
            // implicit-GEMM wgrad: M = O, N = (tap, c), k-block = 32 output pixels; A = g^T and B = shifted x, both MN-major
            const int wt = wg_wt, ht = wg_ht, ng = wg_ng, c0 = wg_c0, u = wg_u, v = wg_v;
#pragma unroll
            for (int c = 0; c < TC_BM / Cfg::EPC; c++)
              tma_load_4d(&tmA, &full_bar[stage], sa + c * (Cfg::BK * 128), m0 + c * Cfg::EPC, wt * p.cv_TW, ht * p.cv_TH, ng * p.cv_TN);
#pragma unroll
            for (int c = 0; c < BN / Cfg::EPC; c++)
              tma_load_4d(&tmB, &full_bar[stage], sb + c * (Cfg::BK * 128), c0 + c * Cfg::EPC, wt * p.cv_TW + v - p.cv_pw,
                          ht * p.cv_TH + u - p.cv_ph, ng * p.cv_TN);
            if (++wg_wt == p.cv_tiles_w) { wg_wt = 0; if (++wg_ht == p.cv_tiles_h) { wg_ht = 0; wg_ng++; } }
          } else if (PAIR) {
            // both CTAs' boxes are counted on the even CTA's barrier (armed there with the pair's total above); B: my half of the rows
            if (!A_MN) {
              tma_load_2d_pair(&tmA, &full_bar[stage], sa, k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < TC_BM / Cfg::EPC; c++) tma_load_2d_pair(&tmA, &full_bar[stage], sa + c * (Cfg::BK * 128), m0 + c * Cfg::EPC, k0);
            }
            if (!B_MN) {
              tma_load_2d_pair(&tmB, &full_bar[stage], sb, k0, n0 + (int)crank * (BN / 2));      // tensor map box: BN / 2 rows
            } else {
              constexpr int HC = BN / Cfg::EPC / 2;
#pragma unroll
              for (int c = 0; c < HC; c++)
                tma_load_2d_pair(&tmB, &full_bar[stage], sb + c * (Cfg::BK * 128), n0 + ((int)crank * HC + c) * Cfg::EPC, k0);
            }
          } else {
          if (!A_MN) {
            tma_load_2d(&tmA, &full_bar[stage], sa, k0, m0);                       // box {BK, 128}: rows = m
          } else {
#pragma unroll
            for (int c = 0; c < TC_BM / Cfg::EPC; c++)                             // box {EPC, BK}: rows = k
              tma_load_2d(&tmA, &full_bar[stage], sa + c * (Cfg::BK * 128), m0 + c * Cfg::EPC, k0);
          }
          if (!B_MN) {
            tma_load_2d(&tmB, &full_bar[stage], sb, k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / Cfg::EPC; c++)
              tma_load_2d(&tmB, &full_bar[stage], sb + c * (Cfg::BK * 128), n0 + c * Cfg::EPC, k0);
          }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (crank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc(KIND, A_MN, B_MN, PAIR ? 2 * TC_BM : TC_BM, BN);
      // K-major: rows 128 B apart, 8-row groups 1024 B apart.  MN-major: 8 k-rows per 1024 B atom (SBO), next 128-byte
      // chunk along MN one region (BK rows x 128 B) further (LBO).
      constexpr uint32_t a_lbo = A_MN ? Cfg::BK * 128 : 16, b_lbo = B_MN ? Cfg::BK * 128 : 16;
      // MN-major tf32 uses the 32-byte-atom swizzle: 4 k-rows (512 B) per atom instead of 8 (1024 B)
      constexpr uint32_t a_lay = (A_MN && KIND == 0) ? 1 : 2, b_lay = (B_MN && KIND == 0) ? 1 : 2;
      constexpr uint32_t a_sbo = (A_MN && KIND == 0) ? 512 : 1024, b_sbo = (B_MN && KIND == 0) ? 512 : 1024;
      constexpr uint32_t a_adv = A_MN ? (Cfg::UMMA_K * 128) >> 4 : 32 >> 4, b_adv = B_MN ? (Cfg::UMMA_K * 128) >> 4 : 32 >> 4;
      // lean issue loop (see umma_bf16_words in tc_ptx.cuh): descriptor low words by one multiply-add per k-block, the high words are
      // compile-time constants, run-time loop parameters are read from shared memory once, no fence per k-block
      constexpr uint32_t a_hi = (a_sbo >> 4) | (1u << 14) | (a_lay << 29), b_hi = (b_sbo >> 4) | (1u << 14) | (b_lay << 29);
      constexpr uint32_t ST16 = Cfg::STAGE_BYTES >> 4;
      const uint32_t a_lo0 = ((smem_u32(smem) & 0x3FFFFu) >> 4) | ((a_lbo >> 4) << 16);
      const uint32_t b_lo0 = (((smem_u32(smem) + Cfg::A_BYTES) & 0x3FFFFu) >> 4) | ((b_lbo >> 4) << 16);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);
      const uint32_t nstages = issue_par[0];
      bool want_ts = issue_par[1] != 0;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int w = wstart; w < total_work; w += wstep) {
        const int split = w / ((p.tiles_m / CS) * p.tiles_n);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait_a(tempty0 + acc * 8, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accum = 0;
#pragma unroll 1
        for (int kb = kb0; kb < kb1; kb++) {
          const uint32_t a_lo = a_lo0 + stage * ST16, b_lo = b_lo0 + stage * ST16;
          mbar_wait_a(full0 + stage * 8, phase);
          if (want_ts) {
            p.dbg[blockIdx.x * 8 + 2] = gtime();
            want_ts = false;
          }
#pragma unroll
          for (int k = 0; k < Cfg::BK / Cfg::UMMA_K; k++)
            umma_words2<KIND, PAIR>(d_tmem, a_lo + (uint32_t)(k * a_adv), a_hi, b_lo + (uint32_t)(k * b_adv), b_hi, idesc, k ? 1u : accum);
          accum = 1;
          if (PAIR) umma_commit_pair_a(empty0 + stage * 8);  // both CTAs' producers
          else umma_commit_a(empty0 + stage * 8);            // smem stage reusable once these MMAs have read it
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        if (PAIR) umma_commit_pair_a(tfull0 + acc * 8);
        else umma_commit_a(tfull0 + acc * 8);                // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue: 8 warps (4..11).  Warp w serves TMEM lanes
    // 32*(w&3)..+31 (hardware lane partition) and every second 32-column chunk (half = (w-4)>>2), so two warps per lane quarter keep
    // twice as many TMEM loads / global stores in flight; everything that does not depend on the accumulator (bias, row -> pixel
    // mapping) is fetched BEFORE waiting for it.
    const int q = warp & 3, half = (warp - 4) >> 2;
    int acc = 0;
    int dbg_tile = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = (p.splits > 1) ? ((p.N & 3) == 0) : ((p.ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(p.D) & 15) == 0);
    const bool vec16_ok = p.D16 && (p.ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(p.D16) & 7) == 0;
    float* stg = stage_base + (warp - 4) * (16 * 36);
    const int c4 = (lane & 7) * 4, rsub = lane >> 3;
    constexpr int NCH = BN / 32;                     // 32-column chunks per tile
    constexpr int MYCH = (NCH + 1) / 2;              // chunks this warp may own
    for (int w = wstart; w < total_work; w += wstep) {
      const int tmg = p.tiles_m / CS;
      const int split = w / (tmg * p.tiles_n);
      const int t = w - split * (tmg * p.tiles_n);
      const int tm = (t % tmg) * CS + (int)crank, tn = t / tmg;
      const int n0 = tn * BN;
      const int m_base = tm * TC_BM + q * 32;
      // ---- accumulator-independent prefetch: bias values of my columns, output row of each of my 8 row slots
      float bz[MYCH][4];
#pragma unroll
      for (int ci = 0; ci < MYCH; ci++) {
        const int nn = n0 + (half + 2 * ci) * 32 + c4;
#pragma unroll
        for (int e = 0; e < 4; e++) bz[ci][e] = (p.splits == 1 && p.bias && half + 2 * ci < NCH && nn + e < p.N) ? __ldg(p.bias + nn + e) : 0.f;
      }
      int rowm[8];
      if (MODE == 1) {
        const int wt = tm % p.cv_tiles_w, r2 = tm / p.cv_tiles_w;
        const int ht = r2 % p.cv_tiles_h, ng = r2 / p.cv_tiles_h;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int tr = q * 32 + 4 * i + rsub;      // tile row -> (image, h, w) of the channels-last output
          const int ww = wt * p.cv_TW + (tr & (p.cv_TW - 1));
          const int hh = ht * p.cv_TH + ((tr >> p.cv_tw_sh) & (p.cv_TH - 1));
          const int ni = ng * p.cv_TN + (tr >> (p.cv_tw_sh + p.cv_th_sh));
          rowm[i] = (ww < p.cv_OW && hh < p.cv_OH && ni < p.cv_NB) ? (ni * p.cv_OH + hh) * p.cv_OW + ww : p.M;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) rowm[i] = m_base + 4 * i + rsub;
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (p.dbg && threadIdx.x == 128) {
        p.dbg[blockIdx.x * 8 + 3] = gtime();
        if (blockIdx.x == 0 && dbg_tile < 24) p.dbg[148 * 8 + 2 * dbg_tile] = gtime();
      }
      // TMEM -> registers -> per-warp smem staging (padded rows, conflict-free 128-bit accesses) -> global: every store instruction
      // writes four full 128-byte row segments.
#pragma unroll
      for (int ci = 0; ci < MYCH; ci++) {
        const int ch = half + 2 * ci;
        if (ch >= NCH) break;
        const int c = ch * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), v);
        const int nn = n0 + c + c4;
        float4 mk[8];
        if (p.mask && p.splits == 1) {               // relu mask of the 8 row slots: all loads in flight while TMEM drains
#pragma unroll
          for (int i = 0; i < 8; i++) {
            mk[i] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (rowm[i] < p.M && nn + 4 <= p.N && vec_ok)
              mk[i] = __ldg(reinterpret_cast<const float4*>(p.mask + (size_t)rowm[i] * (size_t)p.ldd + nn));
          }
        }
        tmem_ld_wait();
        // two passes of 16 rows through an 16 x 36-float staging tile per warp (keeps the 8 staging tiles within the smem budget)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if ((i & 3) == 0) {
            __syncwarp();
            if ((lane >> 4) == (i >> 2)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(stg + (lane & 15) * 36 + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
            __syncwarp();
          }
          const int r = 4 * (i & 3) + rsub;
          const int mm = rowm[i];
          float4 o = *reinterpret_cast<const float4*>(stg + r * 36 + c4);
          if (mm < p.M && nn < p.N) {
            if (p.splits > 1) {
              float* dst = p.ws + ((size_t)split * p.M + mm) * (size_t)p.N + nn;
              if (vec_ok && nn + 4 <= p.N) *reinterpret_cast<float4*>(dst) = o;
              else {
                const float f[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int e = 0; e < 4; e++)
                  if (nn + e < p.N) dst[e] = f[e];
              }
            } else {
              float f[4] = {o.x + bz[ci][0], o.y + bz[ci][1], o.z + bz[ci][2], o.w + bz[ci][3]};
#pragma unroll
              for (int e = 0; e < 4; e++) f[e] = act_apply(f[e], p.act);
              if (p.mask) {
                if (vec_ok && nn + 4 <= p.N) {
                  if (!(mk[i].x > 0.f)) f[0] = 0.f;
                  if (!(mk[i].y > 0.f)) f[1] = 0.f;
                  if (!(mk[i].z > 0.f)) f[2] = 0.f;
                  if (!(mk[i].w > 0.f)) f[3] = 0.f;
                } else {
                  const float* mp = p.mask + (size_t)mm * (size_t)p.ldd + nn;
#pragma unroll
                  for (int e = 0; e < 4; e++)
                    if (nn + e < p.N && !(__ldg(mp + e) > 0.f)) f[e] = 0.f;
                }
              }
              float* dst = p.D + (size_t)mm * (size_t)p.ldd + nn;
              if (vec_ok && nn + 4 <= p.N) *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
              else {
#pragma unroll
                for (int e = 0; e < 4; e++)
                  if (nn + e < p.N) dst[e] = f[e];
              }
              if (p.D16) {
                __nv_bfloat16* d16 = p.D16 + (size_t)mm * (size_t)p.ldd + nn;
                if (vec16_ok && nn + 4 <= p.N) {
                  __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]), hi = __floats2bfloat162_rn(f[2], f[3]);
                  uint2 pk;
                  pk.x = *reinterpret_cast<uint32_t*>(&lo);
                  pk.y = *reinterpret_cast<uint32_t*>(&hi);
                  *reinterpret_cast<uint2*>(d16) = pk;
                } else {
#pragma unroll
                  for (int e = 0; e < 4; e++)
                    if (nn + e < p.N) d16[e] = __float2bfloat16(f[e]);
                }
              }
            }
          }
        }
        __syncwarp();
      }
      if (p.dbg && threadIdx.x == 128 && blockIdx.x == 0 && dbg_tile < 24) p.dbg[148 * 8 + 2 * dbg_tile + 1] = gtime();
      dbg_tile++;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0u));
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (p.splits > 1) {
        // Split-K fix-up without a second kernel.  All CTAs of a launch are co-resident (grid <= #SMs, one work item each
        // when splits > 1), so the `splits` CTAs of a tile meet at a counter, then EACH reduces 1/splits of the tile over
        // all partials in fixed split order (deterministic) and applies bias + activation.
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 128) {
          atomicAdd(p.counters + 2 * t, 1);
          const long long t0 = clock64();
          while (*reinterpret_cast<volatile int*>(p.counters + 2 * t) < p.splits) {
            __nanosleep(64);
            if (clock64() - t0 > 4000000000LL) {
              printf("okl tc kernel: split-K rendezvous timed out (block %d)\n", blockIdx.x);
              __trap();
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        __threadfence();
        {
          const int te = threadIdx.x - 128;           // 0..255
          const int m0 = tm * TC_BM;
          const size_t plane = (size_t)p.M * p.N;
          constexpr int C4 = BN / 4;
          constexpr int TOTAL4 = TC_BM * C4;
          const int chunk = (TOTAL4 + p.splits - 1) / p.splits;
          const int beg = split * chunk, end = min(TOTAL4, beg + chunk);
          for (int idx = beg + te; idx < end; idx += 256) {
            const int r = idx / C4, c = (idx - r * C4) * 4;
            const int mm = m0 + r, nn = n0 + c;
            if (mm >= p.M || nn >= p.N) continue;
            const float* src = p.ws + (size_t)mm * p.N + nn;
            float4 a[4];
#pragma unroll
            for (int u = 0; u < 4; u++) a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            int z = 0;
            for (; z + 3 < p.splits; z += 4) {
              float4 v[4];
#pragma unroll
              for (int u = 0; u < 4; u++) v[u] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(z + u) * plane));
#pragma unroll
              for (int u = 0; u < 4; u++) { a[u].x += v[u].x; a[u].y += v[u].y; a[u].z += v[u].z; a[u].w += v[u].w; }
            }
            for (; z < p.splits; z++) {
              const float4 v0 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)z * plane));
              a[0].x += v0.x; a[0].y += v0.y; a[0].z += v0.z; a[0].w += v0.w;
            }
            float f[4] = {(a[0].x + a[1].x) + (a[2].x + a[3].x), (a[0].y + a[1].y) + (a[2].y + a[3].y),
                          (a[0].z + a[1].z) + (a[2].z + a[3].z), (a[0].w + a[1].w) + (a[2].w + a[3].w)};
#pragma unroll
            for (int j = 0; j < 4; j++) {
              if (p.bias) f[j] += __ldg(p.bias + nn + j);
              f[j] = act_apply(f[j], p.act);
              if (p.mask && !(__ldg(p.mask + (size_t)mm * p.ldd + nn + j) > 0.f)) f[j] = 0.f;
            }
            *reinterpret_cast<float4*>(p.D + (size_t)mm * p.ldd + nn) = make_float4(f[0], f[1], f[2], f[3]);
            if (p.D16) {
#pragma unroll
              for (int j = 0; j < 4; j++) p.D16[(size_t)mm * p.ldd + nn + j] = __float2bfloat16(f[j]);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 128) {
          const int old = atomicAdd(p.counters + 2 * t + 1, 1);
          if (old == p.splits - 1) {          // last CTA to leave: every reader is done, re-arm both counters
            p.counters[2 * t] = 0;
            p.counters[2 * t + 1] = 0;
          }
        }
      }
    }
  }

  if (p.dbg && threadIdx.x == 128) p.dbg[blockIdx.x * 8 + 4] = gtime();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_cg2<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 5] = gtime();
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  size_t n2 = n >> 1;
  for (size_t j = i; j < n2; j += st) {
    float2 v = reinterpret_cast<const float2*>(s)[j];
    reinterpret_cast<__nv_bfloat162*>(d)[j] = __floats2bfloat162_rn(v.x, v.y);
  }
  if (i == 0 && (n & 1)) d[n - 1] = __float2bfloat16(s[n - 1]);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D tensor map over a row-major matrix [outer][inner] with 128-byte swizzle
int make_map_2d(CUtensorMap* map, int kind, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems, uint32_t box_inner,
                uint32_t box_outer, bool atom32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    okl_set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return OKL_ERR_CUDA;
  }
  const int es = kind == 0 ? 4 : 2;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld_elems * es};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                  gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    okl_set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu box=%ux%u)", (int)r,
                  (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems, box_inner, box_outer);
    return OKL_ERR_CUDA;
  }
  return OKL_OK;
}

template <int KIND, bool A_MN, bool B_MN, int BN, int MODE = 0>
int launch_cfg(okl_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const TcGemmParams& p) {
  using Cfg = TcCfg<KIND, BN>;
  auto kern = tc_gemm_kernel<KIND, A_MN, B_MN, BN, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    OKL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  int total = p.tiles_m * p.tiles_n * p.splits;
  int grid = total < ctx->sm_count ? total : ctx->sm_count;
  TcGemmParams q = p;
  const int kb = p.kb_per_split > 0 ? p.kb_per_split : p.kb_total;
  q.stages = total <= grid ? (kb < 2 ? 2 : (kb < Cfg::STAGES ? kb : Cfg::STAGES)) : Cfg::STAGES;   // one tile per CTA: no deeper than its K loop
  const int smem_bytes = Cfg::SMEM_BYTES - (Cfg::STAGES - q.stages) * Cfg::STAGE_BYTES;
  okl_launch(ctx, kern, dim3(grid), dim3(TC_THREADS), smem_bytes, ta, tb, q);
  OKL_LAUNCHED(ctx, "tc_gemm_kernel");
  return OKL_OK;
}

template <bool A_MN, bool B_MN, int BN>
int launch_pair(okl_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const TcGemmParams& p) {
  using Cfg = TcCfg<1, BN>;
  auto kern = tc_gemm_kernel<1, A_MN, B_MN, BN, 0, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    OKL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int units = (p.tiles_m / 2) * p.tiles_n;
  int clusters = ctx->sm_count / 2;
  if (clusters > units) clusters = units;
  TcGemmParams q = p;
  q.stages = Cfg::STAGES;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * 2); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  memset(at, 0, sizeof(at));
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, ta, tb, q);
  OKL_LAUNCHED(ctx, "tc_gemm_kernel");
  return OKL_OK;
}

template <int KIND, bool A_MN, bool B_MN>
int launch_bn(okl_ctx* ctx, int bn, const CUtensorMap& ta, const CUtensorMap& tb, const TcGemmParams& p) {
  switch (bn) {
    case 32: return launch_cfg<KIND, A_MN, B_MN, 32>(ctx, ta, tb, p);
    case 64: return launch_cfg<KIND, A_MN, B_MN, 64>(ctx, ta, tb, p);
    case 128: return launch_cfg<KIND, A_MN, B_MN, 128>(ctx, ta, tb, p);
    case 256: return launch_cfg<KIND, A_MN, B_MN, 256>(ctx, ta, tb, p);
  }
  okl_set_error("unsupported BN %d", bn);
  return OKL_ERR_INVALID;
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

}  // namespace

namespace {
struct Bf16Scratch {
  void* a = nullptr;
  void* b = nullptr;
};
int bf16_operands(okl_ctx* ctx, const float* A, size_t na, const float* B, size_t nb, Bf16Scratch& out);
}  // namespace
int okl_f32_to_bf16(okl_ctx* ctx, const float* src, void* dst, size_t n);

// Generic entry: D[M,N] (row-major, ldd) = act(A . B + bias).
//   a_mn == 0: A stored [M][K] (lda, K contiguous);  a_mn == 1: A stored [K][M] (lda, M contiguous)
//   b_mn == 0: B stored [N][K] (ldb, K contiguous);  b_mn == 1: B stored [K][N] (ldb, N contiguous)
// kind 0 = TF32 (A,B are fp32), kind 1 = BF16 (A,B are bf16).
int okl_tc_gemm_raw(okl_ctx* ctx, int kind, int a_mn, int b_mn, int M, int N, int K, const void* A, long long lda, const void* B,
                    long long ldb, float* D, long long ldd, const float* bias, int act, void* D16, const float* mask = nullptr) {
  const int es = kind == 0 ? 4 : 2;
  const int bk = 128 / es, epc = 128 / es;
  OKL_REQUIRE(M >= 1 && N >= 1 && K >= 1, "tc_gemm: bad shape");
  OKL_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0, "tc_gemm: operands must be 16-byte aligned");
  OKL_REQUIRE((lda * es) % 16 == 0 && (ldb * es) % 16 == 0, "tc_gemm: leading dimensions must be multiples of 16 bytes");
  // tile / split selection
  int bn = env_int("OKL_TC_BN", 0);
  const int tiles_m = (M + TC_BM - 1) / TC_BM;
  const int kb_total = (K + bk - 1) / bk;
  // a GEMM on an auxiliary lane (a weight gradient) runs NEXT TO the main lane's kernel (the matching dX GEMM): it gets a CTA
  // budget that lets both be resident at once instead of queueing for SMs behind it
  const int sm_budget = (ctx->cur_lane != 0 && env_int("OKL_TC_AUX_BUDGET", 1)) ? ctx->sm_count * 42 / 100 : ctx->sm_count;
  if (bn == 0 && sm_budget < ctx->sm_count) {
    auto tiles = [&](int b) { return (long long)tiles_m * ((N + b - 1) / b); };
    for (int b = 64; b <= 256 && bn == 0; b *= 2)             // smallest tile (most CTAs) that stays within the budget
      if (tiles(b) <= sm_budget && (b == 64 || N > b / 2)) bn = b;
  }
  if (bn == 0) {
    auto tiles = [&](int b) { return (long long)tiles_m * ((N + b - 1) / b); };
    if (N > 128 && tiles(256) >= ctx->sm_count) bn = 256;
    else if (N > 64 && tiles(128) >= ctx->sm_count / 2) bn = 128;
    else if (N > 32 && (tiles(64) >= ctx->sm_count / 2 || N > 64)) bn = N > 128 ? 128 : 64;
    else bn = N > 32 ? 64 : 32;
  }
  if (b_mn && bn < epc) bn = epc;                            // an MN-major B tile is loaded in 128-byte chunks along N
  // keep every TMA box inside the tensor's inner extent (tiny operands are the FFMA kernels' job)
  if ((a_mn ? M : K) < (a_mn ? epc : bk) || (b_mn ? N : K) < (b_mn ? epc : bk)) {
    okl_set_error("tc_gemm: operand smaller than one 128-byte TMA box (M=%d N=%d K=%d)", M, N, K);
    return OKL_ERR_UNSUPPORTED;
  }
  const int tiles_n = (N + bn - 1) / bn;
  int splits = env_int("OKL_TC_SPLITS", 0);
  {
    // split-K only when every (tile, split) work item gets its own co-resident CTA (single round) and the fix-up can use
    // 128-bit accesses
    const long long tiles = (long long)tiles_m * tiles_n;
    const bool can_split = (N & 3) == 0 && (ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0 && tiles <= 2048;
    int max_splits = (int)(sm_budget / tiles);
    if (max_splits > kb_total / 4) max_splits = kb_total / 4;
    if (max_splits > 16) max_splits = 16;                  // measured (256x5408x128): 16 splits 15.6 us, 29 splits 17.4 us - the fix-up grows with the split count
    if (max_splits < 1 || !can_split) max_splits = 1;
    if (splits == 0) splits = max_splits;
    if (splits > max_splits) splits = max_splits;
  }
  int kb_per_split = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per_split - 1) / kb_per_split;

  CUtensorMap ta, tb;
  int s;
  if (!a_mn) s = make_map_2d(&ta, kind, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, bk, TC_BM);
  else s = make_map_2d(&ta, kind, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, epc, bk, kind == 0);
  if (s != OKL_OK) return s;
  // CTA pairs (cta_group::2) for the large bf16 products: full tiles only, an even number of M tiles, enough pairs to fill the chip
  const bool pair = kind == 1 && splits == 1 && (bn == 128 || bn == 256) && tiles_m % 2 == 0 && N % bn == 0 && M % TC_BM == 0 &&
                    (long long)(tiles_m / 2) * tiles_n >= ctx->sm_count / 2 && ctx->cur_lane == 0 && env_int("OKL_TC_PAIR", 1);
  if (!b_mn) s = make_map_2d(&tb, kind, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, bk, pair ? bn / 2 : bn);
  else s = make_map_2d(&tb, kind, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, epc, bk, kind == 0);
  if (s != OKL_OK) return s;

  TcGemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.tiles_m = tiles_m; p.tiles_n = tiles_n; p.splits = splits; p.kb_per_split = kb_per_split; p.kb_total = kb_total;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.tiles_m = tiles_m; p.tiles_n = tiles_n; p.splits = splits; p.kb_per_split = kb_per_split; p.kb_total = kb_total;
  p.D = D; p.ldd = ldd; p.bias = bias; p.act = act; p.ws = nullptr; p.D16 = (__nv_bfloat16*)D16; p.mask = mask;
  p.counters = nullptr;
  p.dbg = (unsigned long long*)ctx->tc_dbg;
  if (splits > 1) {
    s = okl_ensure_workspace(ctx, (size_t)splits * M * N * sizeof(float));
    if (s != OKL_OK) return s;
    p.ws = (float*)ctx->workspace;
    if (!ctx->tile_counters) {
      OKL_CHECK_CUDA(cudaMalloc(&ctx->tile_counters, 4096 * sizeof(int)));
      OKL_CHECK_CUDA(cudaMemsetAsync(ctx->tile_counters, 0, 4096 * sizeof(int), ctx->stream));
    }
    p.counters = ctx->tile_counters;
  }
  if (pair) {
    if (bn == 256) {
      if (!a_mn && !b_mn) return launch_pair<false, false, 256>(ctx, ta, tb, p);
      if (!a_mn && b_mn) return launch_pair<false, true, 256>(ctx, ta, tb, p);
      if (a_mn && !b_mn) return launch_pair<true, false, 256>(ctx, ta, tb, p);
      return launch_pair<true, true, 256>(ctx, ta, tb, p);
    }
    if (!a_mn && !b_mn) return launch_pair<false, false, 128>(ctx, ta, tb, p);
    if (!a_mn && b_mn) return launch_pair<false, true, 128>(ctx, ta, tb, p);
    if (a_mn && !b_mn) return launch_pair<true, false, 128>(ctx, ta, tb, p);
    return launch_pair<true, true, 128>(ctx, ta, tb, p);
  }
  if (kind == 0) {
    if (!a_mn && !b_mn) s = launch_bn<0, false, false>(ctx, bn, ta, tb, p);
    else if (!a_mn && b_mn) s = launch_bn<0, false, true>(ctx, bn, ta, tb, p);
    else if (a_mn && !b_mn) s = launch_bn<0, true, false>(ctx, bn, ta, tb, p);
    else s = launch_bn<0, true, true>(ctx, bn, ta, tb, p);
  } else {
    if (!a_mn && !b_mn) s = launch_bn<1, false, false>(ctx, bn, ta, tb, p);
    else if (!a_mn && b_mn) s = launch_bn<1, false, true>(ctx, bn, ta, tb, p);
    else if (a_mn && !b_mn) s = launch_bn<1, true, false>(ctx, bn, ta, tb, p);
    else s = launch_bn<1, true, true>(ctx, bn, ta, tb, p);
  }
  return s;
}

extern "C" int okl_tc_gemm(okl_ctx* ctx, int kind, int a_mn, int b_mn, int M, int N, int K, const void* A, long long lda, const void* B,
                           long long ldb, void* D, long long ldd, const void* bias, int act) {
  OKL_REQUIRE(ctx && A && B && D, "NULL argument");
  OKL_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (tf32) or 1 (bf16)");
  OKL_REQUIRE(ctx->cc >= 100, "tcgen05 kernels need compute capability 10.x");
  if (kind == 0)
    return okl_tc_gemm_raw(ctx, 0, a_mn, b_mn, M, N, K, A, lda, B, ldb, (float*)D, ldd, (const float*)bias, act, nullptr);
  // bf16: convert the fp32 operands (dense, ld == inner extent) into scratch first
  size_t na = (size_t)(a_mn ? K : M) * lda, nb = (size_t)(b_mn ? K : N) * ldb;
  Bf16Scratch sc;
  int s = bf16_operands(ctx, (const float*)A, na, (const float*)B, nb, sc);
  if (s != OKL_OK) return s;
  return okl_tc_gemm_raw(ctx, 1, a_mn, b_mn, M, N, K, sc.a, lda, sc.b, ldb, (float*)D, ldd, (const float*)bias, act, nullptr);
}

// debug hook (not in okl.h): per-CTA phase timestamps of the next tcgen05 GEMM launches are written to `buf` (8 u64 per CTA)
extern "C" int okl_debug_tc_timestamps(okl_ctx* ctx, void* buf) {
  ctx->tc_dbg = buf;
  return OKL_OK;
}

int okl_f32_to_bf16(okl_ctx* ctx, const float* src, void* dst, size_t n) {
  if (n == 0) return OKL_OK;
  f32_to_bf16_kernel<<<okl_grid_1d(ctx, (n + 1) / 2, 256), 256, 0, ctx->stream>>>(src, (__nv_bfloat16*)dst, n);
  OKL_LAUNCHED(ctx, "f32_to_bf16");
  return OKL_OK;
}


// ================================================================================================ implicit-GEMM convolution
namespace {

typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 4-D fp32 tensor map over a channels-last activation tensor (N, H, W, C): dims {C, W, H, N}
int make_map_nhwc(CUtensorMap* map, const void* base, int C, int W, int H, int N, int box_c, int box_w, int box_h, int box_n, bool atom32,
                  int kind = 0) {
  const cuuint64_t es = kind == 0 ? 4 : 2;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    okl_set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return OKL_ERR_CUDA;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstride[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim,
                  gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    okl_set_error("cuTensorMapEncodeTiled(4d) failed with CUresult %d (C=%d W=%d H=%d N=%d box=%d,%d,%d,%d)", (int)r, C, W, H, N, box_c,
                  box_w, box_h, box_n);
    return OKL_ERR_CUDA;
  }
  return OKL_OK;
}

int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

// pixels-per-tile rectangle TW x TH x TN (all powers of two, product = `pixels`) for an OH x OW image
void pick_tile(int OH, int OW, int pixels, int& TW, int& TH, int& TN) {
  TW = pow2_ceil(OW); if (TW > pixels) TW = pixels;
  TH = pow2_ceil(OH); if (TH > pixels / TW) TH = pixels / TW;
  TN = pixels / (TW * TH);
}

// W2[o][tap][c] = W[o][c][tap]  (K-major filter bank for the channels-last implicit GEMM)
// Wf[c][tap][o] = W[o][c][T-1-tap]  (flipped + transposed bank: dgrad as a convolution of g)
__global__ void flip_filters_kernel(const float* __restrict__ W, float* __restrict__ Wf, int O, int C, int T) {
  const size_t total = (size_t)O * C * T;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % O);
    const size_t r = i / O;
    const int t = (int)(r % T), c = (int)(r / T);
    Wf[i] = W[((size_t)o * C + c) * T + (T - 1 - t)];
  }
}

struct ConvShape { int N, C, O, H, W, KH, KW, PH, PW, OH, OW; };

int conv_shape(const okl_conv_desc* d, ConvShape& s) {
  s.N = d->N; s.C = d->C; s.O = d->O;
  if (d->nd == 1) { s.H = 1; s.W = d->in[0]; s.KH = 1; s.KW = d->k[0]; s.PH = 0; s.PW = d->p[0]; }
  else { s.H = d->in[0]; s.W = d->in[1]; s.KH = d->k[0]; s.KW = d->k[1]; s.PH = d->p[0]; s.PW = d->p[1]; }
  s.OH = s.H + 2 * s.PH - s.KH + 1;
  s.OW = s.W + 2 * s.PW - s.KW + 1;
  return OKL_OK;
}

int pick_bn(int n) { return n >= 256 && n % 256 == 0 ? 256 : (n >= 128 && n % 128 == 0 ? 128 : (n >= 64 && n % 64 == 0 ? 64 : 32)); }

// y[N,OH,OW,Oc] = act(conv(x[N,H,W,Cc], W2[Oc][T*Cc]) + bias)  with padding (ph, pw), stride 1, everything channels-last
int conv_igemm_fprop(okl_ctx* ctx, const void* x, int N, int H, int W, int Cc, const void* W2, int Oc, int KH, int KW, int ph, int pw,
                     int OH, int OW, const float* bias, int act, float* y, const float* mask = nullptr, int kind = 0, void* y16 = nullptr) {
  const int T = KH * KW;
  const int bkc = kind == 0 ? 32 : 64;                  // channels per k-block (one 128-byte swizzle row)
  int TW, TH, TN;
  pick_tile(OH, OW, TC_BM, TW, TH, TN);
  int bn = Oc >= 256 ? 256 : (Oc > 64 ? 128 : (Oc > 32 ? 64 : 32));
  const int tiles_w = (OW + TW - 1) / TW, tiles_h = (OH + TH - 1) / TH, tiles_ng = (N + TN - 1) / TN;
  long long tiles_m = (long long)tiles_w * tiles_h * tiles_ng;
  while (bn > 64 && tiles_m * ((Oc + bn - 1) / bn) < ctx->sm_count) bn >>= 1;     // more, smaller tiles when the grid would not fill the chip
  CUtensorMap ta, tb;
  int s = make_map_nhwc(&ta, x, Cc, W, H, N, bkc, TW, TH, TN, false, kind);
  if (s != OKL_OK) return s;
  s = make_map_2d(&tb, kind, W2, (uint64_t)T * Cc, (uint64_t)Oc, (uint64_t)T * Cc, bkc, bn);
  if (s != OKL_OK) return s;
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = N * OH * OW; p.N = Oc; p.K = T * Cc;
  p.tiles_m = (int)tiles_m; p.tiles_n = (Oc + bn - 1) / bn; p.splits = 1;
  p.kb_total = T * (Cc / bkc); p.kb_per_split = p.kb_total;
  p.D = y; p.ldd = Oc; p.bias = bias; p.act = act; p.mask = mask; p.D16 = (__nv_bfloat16*)y16;
  p.cv_TW = TW; p.cv_TH = TH; p.cv_TN = TN; p.cv_tw_sh = ilog2(TW); p.cv_th_sh = ilog2(TH);
  p.cv_tiles_w = tiles_w; p.cv_tiles_h = tiles_h; p.cv_OH = OH; p.cv_OW = OW; p.cv_NB = N;
  p.cv_C = Cc; p.cv_cblocks = Cc / bkc; p.cv_kw = KW; p.cv_ph = ph; p.cv_pw = pw;
  p.dbg = (unsigned long long*)ctx->tc_dbg;
  if (kind == 1) {
    switch (bn) {
      case 32: return launch_cfg<1, false, false, 32, 1>(ctx, ta, tb, p);
      case 64: return launch_cfg<1, false, false, 64, 1>(ctx, ta, tb, p);
      case 128: return launch_cfg<1, false, false, 128, 1>(ctx, ta, tb, p);
      default: return launch_cfg<1, false, false, 256, 1>(ctx, ta, tb, p);
    }
  }
  switch (bn) {
    case 32: return launch_cfg<0, false, false, 32, 1>(ctx, ta, tb, p);
    case 64: return launch_cfg<0, false, false, 64, 1>(ctx, ta, tb, p);
    case 128: return launch_cfg<0, false, false, 128, 1>(ctx, ta, tb, p);
    default: return launch_cfg<0, false, false, 256, 1>(ctx, ta, tb, p);
  }
}

struct Scratch {            // pool block released back when the enqueueing call returns (stream-ordered reuse)
  okl_ctx* ctx; void* p = nullptr;
  Scratch(okl_ctx* c) : ctx(c) {}
  int alloc(size_t bytes) { return okl_malloc(ctx, bytes, &p); }
  ~Scratch() { if (p) okl_free(ctx, p); }
};

}  // namespace

bool okl_tc_conv_supported(okl_ctx* ctx, const okl_conv_desc* d) {
  if (ctx->cc < 100 || env_int("OKL_TC_DISABLE", 0) || env_int("OKL_TC_CONV_DISABLE", 0)) return false;
  if (d->nd > 2 || d->x_layout != OKL_NXC || d->y_layout != OKL_NXC) return false;
  for (int i = 0; i < d->nd; i++)
    if (d->s[i] != 1) return false;
  if (d->C % 32 || d->O % 32) return false;             // 32 fp32 channels = one 128-byte swizzle row, for x (fprop) and g (dgrad)
  ConvShape s;
  conv_shape(d, s);
  if (s.OH < 1 || s.OW < 1 || s.KH > 16 || s.KW > 16) return false;
  if (s.PH > s.KH - 1 || s.PW > s.KW - 1) return false;  // dgrad runs as a convolution with padding k-1-p >= 0
  return true;
}

int okl_tc_conv_fprop(okl_ctx* ctx, const okl_conv_desc* d, const float* x, const float* W, const float* b, float* y) {
  ConvShape s;
  conv_shape(d, s);
  const int T = s.KH * s.KW;
  Scratch w2(ctx);
  int st = w2.alloc((size_t)s.O * T * s.C * sizeof(float));
  if (st != OKL_OK) return st;
  st = okl_permute(ctx, w2.p, W, s.O, s.C, T, OKL_NCX);               // (O, C, T) -> (O, T, C)
  if (st != OKL_OK) return st;
  return conv_igemm_fprop(ctx, x, s.N, s.H, s.W, s.C, (const float*)w2.p, s.O, s.KH, s.KW, s.PH, s.PW, s.OH, s.OW, b, d->act, y);
}

int okl_tc_conv_dgrad(okl_ctx* ctx, const okl_conv_desc* d, const float* g, const float* W, float* dx, const float* mask) {
  // dx = conv(g, flip(W)^T) with padding k-1-p: the same implicit GEMM with the roles of C and O exchanged
  ConvShape s;
  conv_shape(d, s);
  const int T = s.KH * s.KW;
  Scratch wf(ctx);
  int st = wf.alloc((size_t)s.O * T * s.C * sizeof(float));
  if (st != OKL_OK) return st;
  flip_filters_kernel<<<okl_grid_1d(ctx, (size_t)s.O * T * s.C, 256), 256, 0, ctx->stream>>>(W, (float*)wf.p, s.O, s.C, T);
  OKL_LAUNCHED(ctx, "flip_filters");
  return conv_igemm_fprop(ctx, g, s.N, s.OH, s.OW, s.O, (const float*)wf.p, s.C, s.KH, s.KW, s.KH - 1 - s.PH, s.KW - 1 - s.PW, s.H, s.W,
                          nullptr, OKL_ACT_NONE, dx, mask);
}

int okl_tc_conv_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const float* x, const float* g, float* dW) {
  ConvShape s;
  conv_shape(d, s);
  const int T = s.KH * s.KW;
  const int M = s.O, N = T * s.C;
  int TW, TH, TN;
  pick_tile(s.OH, s.OW, 32, TW, TH, TN);
  const int tiles_w = (s.OW + TW - 1) / TW, tiles_h = (s.OH + TH - 1) / TH, tiles_ng = (s.N + TN - 1) / TN;
  const int kb_total = tiles_w * tiles_h * tiles_ng;
  const int bn = pick_bn(s.C);                                          // an N tile never straddles two taps
  const int tiles_m = (M + TC_BM - 1) / TC_BM, tiles_n = N / bn;
  Scratch dw2(ctx);
  int st = dw2.alloc((size_t)M * N * sizeof(float));
  if (st != OKL_OK) return st;
  CUtensorMap ta, tb;
  st = make_map_nhwc(&ta, g, s.O, s.OW, s.OH, s.N, 32, TW, TH, TN, true);
  if (st != OKL_OK) return st;
  st = make_map_nhwc(&tb, x, s.C, s.W, s.H, s.N, 32, TW, TH, TN, true);
  if (st != OKL_OK) return st;
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = kb_total * 32;
  p.tiles_m = tiles_m; p.tiles_n = tiles_n;
  long long tiles = (long long)tiles_m * tiles_n;
  int splits = (int)(ctx->sm_count / tiles);
  if (splits > kb_total / 4) splits = kb_total / 4;
  if (splits > 32) splits = 32;
  if (splits < 1) splits = 1;
  p.kb_total = kb_total; p.kb_per_split = (kb_total + splits - 1) / splits;
  p.splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.D = (float*)dw2.p; p.ldd = N; p.bias = nullptr; p.act = OKL_ACT_NONE;
  if (p.splits > 1) {
    st = okl_ensure_workspace(ctx, (size_t)p.splits * M * N * sizeof(float));
    if (st != OKL_OK) return st;
    p.ws = (float*)ctx->workspace;
    if (!ctx->tile_counters) {
      OKL_CHECK_CUDA(cudaMalloc(&ctx->tile_counters, 4096 * sizeof(int)));
      OKL_CHECK_CUDA(cudaMemsetAsync(ctx->tile_counters, 0, 4096 * sizeof(int), ctx->stream));
    }
    OKL_REQUIRE(tiles <= 2048, "conv wgrad: too many tiles for split-K");
    p.counters = ctx->tile_counters;
  }
  p.cv_TW = TW; p.cv_TH = TH; p.cv_TN = TN; p.cv_tw_sh = ilog2(TW); p.cv_th_sh = ilog2(TH);
  p.cv_tiles_w = tiles_w; p.cv_tiles_h = tiles_h; p.cv_OH = s.OH; p.cv_OW = s.OW; p.cv_NB = s.N;
  p.cv_C = s.C; p.cv_cblocks = s.C / 32; p.cv_kw = s.KW; p.cv_ph = s.PH; p.cv_pw = s.PW;
  switch (bn) {
    case 32: st = launch_cfg<0, true, true, 32, 2>(ctx, ta, tb, p); break;
    case 64: st = launch_cfg<0, true, true, 64, 2>(ctx, ta, tb, p); break;
    case 128: st = launch_cfg<0, true, true, 128, 2>(ctx, ta, tb, p); break;
    default: st = launch_cfg<0, true, true, 256, 2>(ctx, ta, tb, p); break;
  }
  if (st != OKL_OK) return st;
  return okl_permute(ctx, dW, dw2.p, s.O, s.C, T, OKL_NXC);             // (O, T, C) -> (O, C, T)
}

// ------------------------------------------------------------------------------------------------ dense entry points
bool okl_tc_dense_supported(okl_ctx* ctx, int M, int K, int N) {
  if (ctx->cc < 100) return false;
  if (env_int("OKL_TC_DISABLE", 0)) return false;
  // TMA needs 16-byte multiples for every leading dimension used: X ld = K, W ld = N, G ld = N (fp32: multiples of 4;
  // bf16 shadows: multiples of 8)
  const int q = ctx->precision == OKL_BF16 ? 8 : 4;
  if (K % q || N % q) return false;
  if (M < 64 || K < 64 || N < 64) return false;              // every TMA box must fit inside its operand
  if ((long long)M * K * N < (1LL << 18)) return false;     // tiny products: launch-bound either way, keep FFMA
  return true;
}

namespace {
// BF16 mode, first cut: operands are converted into a scratch area per call (the fp32 tensors stay the storage of record).
int bf16_operands(okl_ctx* ctx, const float* A, size_t na, const float* B, size_t nb, Bf16Scratch& out) {
  size_t abytes = ((na * 2 + 255) / 256) * 256, bbytes = ((nb * 2 + 255) / 256) * 256;
  if (ctx->bf16_scratch_bytes < abytes + bbytes) {
    if (ctx->capturing) {
      okl_set_error("bf16 scratch must be sized before graph capture (run one eager step first)");
      return OKL_ERR_INVALID;
    }
    OKL_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->bf16_scratch) cudaFree(ctx->bf16_scratch);
    OKL_CHECK_CUDA(cudaMalloc(&ctx->bf16_scratch, abytes + bbytes));
    ctx->bf16_scratch_bytes = abytes + bbytes;
  }
  out.a = ctx->bf16_scratch;
  out.b = (uint8_t*)ctx->bf16_scratch + abytes;
  int s = okl_f32_to_bf16(ctx, A, out.a, na);
  if (s != OKL_OK) return s;
  return okl_f32_to_bf16(ctx, B, out.b, nb);
}
}  // namespace

int okl_tc_dense_fwd(okl_ctx* ctx, const float* X, const float* W, const float* b, float* Y, int M, int K, int N, int act) {
  if (ctx->precision == OKL_BF16) {
    Bf16Scratch sc;
    int s = bf16_operands(ctx, X, (size_t)M * K, W, (size_t)K * N, sc);
    if (s != OKL_OK) return s;
    return okl_tc_gemm_raw(ctx, 1, 0, 1, M, N, K, sc.a, K, sc.b, N, Y, N, b, act, nullptr);
  }
  return okl_tc_gemm_raw(ctx, 0, 0, 1, M, N, K, X, K, W, N, Y, N, b, act, nullptr);
}

int okl_tc_dense_dx(okl_ctx* ctx, const float* G, const float* W, float* dX, int M, int K, int N, const float* mask) {
  // dX[M,K] = G[M,N] . W^T : reduction over N; B'(n'=kin, k'=nout) = W[kin][nout] is K-major as stored
  if (ctx->precision == OKL_BF16) {
    Bf16Scratch sc;
    int s = bf16_operands(ctx, G, (size_t)M * N, W, (size_t)K * N, sc);
    if (s != OKL_OK) return s;
    return okl_tc_gemm_raw(ctx, 1, 0, 0, M, K, N, sc.a, N, sc.b, N, dX, K, nullptr, OKL_ACT_NONE, nullptr, mask);
  }
  return okl_tc_gemm_raw(ctx, 0, 0, 0, M, K, N, G, N, W, N, dX, K, nullptr, OKL_ACT_NONE, nullptr, mask);
}

int okl_tc_dense_dw(okl_ctx* ctx, const float* X, const float* G, float* dW, int M, int K, int N) {
  // dW[K,N] = X^T . G : reduction over the batch M; both operands MN-major as stored
  if (ctx->precision == OKL_BF16) {
    Bf16Scratch sc;
    int s = bf16_operands(ctx, X, (size_t)M * K, G, (size_t)M * N, sc);
    if (s != OKL_OK) return s;
    return okl_tc_gemm_raw(ctx, 1, 1, 1, K, N, M, sc.a, K, sc.b, N, dW, N, nullptr, OKL_ACT_NONE, nullptr);
  }
  return okl_tc_gemm_raw(ctx, 0, 1, 1, K, N, M, X, K, G, N, dW, N, nullptr, OKL_ACT_NONE, nullptr);
}

// ------------------------------------------------------------------------------------------------ BF16 operands kept by the caller
// The Python layers keep a bf16 shadow next to every fp32 activation / weight tensor that feeds a GEMM (written by the producing
// kernel's epilogue), so no operand is converted per call.
extern "C" int okl_to_bf16(okl_ctx* ctx, const void* src_f32, void* dst_bf16, size_t n) {
  OKL_REQUIRE(ctx && (n == 0 || (src_f32 && dst_bf16)), "NULL argument");
  return okl_f32_to_bf16(ctx, (const float*)src_f32, dst_bf16, n);
}

extern "C" int okl_dense_fwd_bf16(okl_ctx* ctx, const void* X16, const void* W16, const void* b, void* Y, void* Y16, int M, int K, int N,
                                  int act) {
  OKL_REQUIRE(ctx && X16 && W16 && Y, "NULL argument");
  OKL_REQUIRE(M >= 1 && K >= 64 && N >= 64 && K % 8 == 0 && N % 8 == 0, "okl_dense_fwd_bf16: unsupported shape M=%d K=%d N=%d", M, K, N);
  OKL_REQUIRE(act >= OKL_ACT_NONE && act <= OKL_ACT_TANH, "bad activation %d", act);
  return okl_tc_gemm_raw(ctx, 1, 0, 1, M, N, K, X16, K, W16, N, (float*)Y, N, (const float*)b, act, Y16);
}

// dW = X^T G, db = sum_rows G (from the fp32 G), dX = (G W^T) (* relu mask), dX16 = bf16 shadow of dX.  Any output may be NULL.
extern "C" int okl_dense_bwd_bf16(okl_ctx* ctx, const void* X16, const void* W16, const void* G16, const void* G, void* dX, void* dX16,
                                  void* dW, void* db, int M, int K, int N, const void* dx_relu_mask) {
  OKL_REQUIRE(ctx && X16 && W16 && G16, "NULL argument");
  OKL_REQUIRE(M >= 64 && K >= 64 && N >= 64 && K % 8 == 0 && N % 8 == 0 && M % 8 == 0, "okl_dense_bwd_bf16: unsupported shape M=%d K=%d N=%d", M,
              K, N);
  int s;
  if (dW) {
    s = okl_tc_gemm_raw(ctx, 1, 1, 1, K, N, M, X16, K, G16, N, (float*)dW, N, nullptr, OKL_ACT_NONE, nullptr);
    if (s != OKL_OK) return s;
  }
  if (db) {
    OKL_REQUIRE(G, "okl_dense_bwd_bf16: the fp32 gradient is needed for db");
    s = okl_colsum(ctx, (const float*)G, (float*)db, M, N);
    if (s != OKL_OK) return s;
  }
  if (dX) {
    s = okl_tc_gemm_raw(ctx, 1, 0, 0, M, K, N, G16, N, W16, N, (float*)dX, K, nullptr, OKL_ACT_NONE, dX16, (const float*)dx_relu_mask);
    if (s != OKL_OK) return s;
  }
  return OKL_OK;
}

// ------------------------------------------------------------------------------------------------ BF16 implicit-GEMM convolution
namespace {
// W2_16[o][tap][c] = bf16(W[o][c][tap])  and  Wf_16[c][tap][o] = bf16(W[o][c][T-1-tap])
__global__ void filters_to_bf16_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int O, int C, int T, int flip) {
  const size_t total = (size_t)O * C * T;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    if (!flip) {
      const int c = (int)(i % C);
      const size_t r = i / C;
      const int t = (int)(r % T), o = (int)(r / T);
      out[i] = __float2bfloat16(W[((size_t)o * C + c) * T + t]);
    } else {
      const int o = (int)(i % O);
      const size_t r = i / O;
      const int t = (int)(r % T), c = (int)(r / T);
      out[i] = __float2bfloat16(W[((size_t)o * C + c) * T + (T - 1 - t)]);
    }
  }
}
}  // namespace

extern "C" int okl_conv_bf16_supported(okl_ctx* ctx, const okl_conv_desc* d) {
  if (!ctx || !d || !okl_tc_conv_supported(ctx, d)) return 0;
  return (d->C % 64 == 0 && d->O % 64 == 0) ? 1 : 0;
}

// x16 / y16 / g16 / dx16 are channels-last bf16 shadows of the fp32 tensors; filters stay fp32 masters (converted per call: tiny)
extern "C" int okl_conv_fprop_bf16(okl_ctx* ctx, const okl_conv_desc* d, const void* x16, const void* W, const void* b, void* y, void* y16) {
  OKL_REQUIRE(ctx && d && x16 && W && y, "NULL argument");
  OKL_REQUIRE(okl_conv_bf16_supported(ctx, d), "okl_conv_fprop_bf16: unsupported shape / layout");
  ConvShape s;
  conv_shape(d, s);
  const int T = s.KH * s.KW;
  Scratch w2(ctx);
  int st = w2.alloc((size_t)s.O * T * s.C * 2);
  if (st != OKL_OK) return st;
  filters_to_bf16_kernel<<<okl_grid_1d(ctx, (size_t)s.O * T * s.C, 256), 256, 0, ctx->stream>>>((const float*)W, (__nv_bfloat16*)w2.p, s.O, s.C, T, 0);
  OKL_LAUNCHED(ctx, "filters_to_bf16");
  return conv_igemm_fprop(ctx, x16, s.N, s.H, s.W, s.C, w2.p, s.O, s.KH, s.KW, s.PH, s.PW, s.OH, s.OW, (const float*)b, d->act, (float*)y,
                          nullptr, 1, y16);
}

extern "C" int okl_conv_dgrad_bf16(okl_ctx* ctx, const okl_conv_desc* d, const void* g16, const void* W, void* dx, void* dx16,
                                   const void* dx_relu_mask) {
  OKL_REQUIRE(ctx && d && g16 && W && dx, "NULL argument");
  OKL_REQUIRE(okl_conv_bf16_supported(ctx, d), "okl_conv_dgrad_bf16: unsupported shape / layout");
  ConvShape s;
  conv_shape(d, s);
  const int T = s.KH * s.KW;
  Scratch wf(ctx);
  int st = wf.alloc((size_t)s.O * T * s.C * 2);
  if (st != OKL_OK) return st;
  filters_to_bf16_kernel<<<okl_grid_1d(ctx, (size_t)s.O * T * s.C, 256), 256, 0, ctx->stream>>>((const float*)W, (__nv_bfloat16*)wf.p, s.O, s.C, T, 1);
  OKL_LAUNCHED(ctx, "filters_to_bf16");
  return conv_igemm_fprop(ctx, g16, s.N, s.OH, s.OW, s.O, wf.p, s.C, s.KH, s.KW, s.KH - 1 - s.PH, s.KW - 1 - s.PW, s.H, s.W, nullptr,
                          OKL_ACT_NONE, (float*)dx, (const float*)dx_relu_mask, 1, dx16);
}

extern "C" int okl_conv_wgrad_bf16(okl_ctx* ctx, const okl_conv_desc* d, const void* x16, const void* g16, const void* g, void* dW, void* db) {
  OKL_REQUIRE(ctx && d && x16 && g16, "NULL argument");
  OKL_REQUIRE(okl_conv_bf16_supported(ctx, d), "okl_conv_wgrad_bf16: unsupported shape / layout");
  ConvShape s;
  conv_shape(d, s);
  const int T = s.KH * s.KW;
  int st;
  if (dW) {
    const int M = s.O, N = T * s.C;
    int TW, TH, TN;
    pick_tile(s.OH, s.OW, 64, TW, TH, TN);                              // 64 pixels = one bf16 k-block
    const int tiles_w = (s.OW + TW - 1) / TW, tiles_h = (s.OH + TH - 1) / TH, tiles_ng = (s.N + TN - 1) / TN;
    const int kb_total = tiles_w * tiles_h * tiles_ng;
    const int bn = pick_bn(s.C) < 64 ? 64 : pick_bn(s.C);
    const int tiles_m = (M + TC_BM - 1) / TC_BM, tiles_n = N / bn;
    Scratch dw2(ctx);
    st = dw2.alloc((size_t)M * N * sizeof(float));
    if (st != OKL_OK) return st;
    CUtensorMap ta, tb;
    st = make_map_nhwc(&ta, g16, s.O, s.OW, s.OH, s.N, 64, TW, TH, TN, false, 1);
    if (st != OKL_OK) return st;
    st = make_map_nhwc(&tb, x16, s.C, s.W, s.H, s.N, 64, TW, TH, TN, false, 1);
    if (st != OKL_OK) return st;
    TcGemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = kb_total * 64;
    p.tiles_m = tiles_m; p.tiles_n = tiles_n;
    long long tiles = (long long)tiles_m * tiles_n;
    int splits = (int)(ctx->sm_count / tiles);
    if (splits > kb_total / 4) splits = kb_total / 4;
    if (splits > 32) splits = 32;
    if (splits < 1) splits = 1;
    p.kb_total = kb_total; p.kb_per_split = (kb_total + splits - 1) / splits;
    p.splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.D = (float*)dw2.p; p.ldd = N;
    if (p.splits > 1) {
      st = okl_ensure_workspace(ctx, (size_t)p.splits * M * N * sizeof(float));
      if (st != OKL_OK) return st;
      p.ws = (float*)ctx->workspace;
      if (!ctx->tile_counters) {
        OKL_CHECK_CUDA(cudaMalloc(&ctx->tile_counters, 4096 * sizeof(int)));
        OKL_CHECK_CUDA(cudaMemsetAsync(ctx->tile_counters, 0, 4096 * sizeof(int), ctx->stream));
      }
      OKL_REQUIRE(tiles <= 2048, "conv wgrad: too many tiles for split-K");
      p.counters = ctx->tile_counters;
    }
    p.cv_TW = TW; p.cv_TH = TH; p.cv_TN = TN; p.cv_tw_sh = ilog2(TW); p.cv_th_sh = ilog2(TH);
    p.cv_tiles_w = tiles_w; p.cv_tiles_h = tiles_h; p.cv_OH = s.OH; p.cv_OW = s.OW; p.cv_NB = s.N;
    p.cv_C = s.C; p.cv_cblocks = s.C / 64; p.cv_kw = s.KW; p.cv_ph = s.PH; p.cv_pw = s.PW;
    switch (bn) {
      case 64: st = launch_cfg<1, true, true, 64, 2>(ctx, ta, tb, p); break;
      case 128: st = launch_cfg<1, true, true, 128, 2>(ctx, ta, tb, p); break;
      default: st = launch_cfg<1, true, true, 256, 2>(ctx, ta, tb, p); break;
    }
    if (st != OKL_OK) return st;
    st = okl_permute(ctx, dW, dw2.p, s.O, s.C, T, OKL_NXC);             // (O, T, C) -> (O, C, T)
    if (st != OKL_OK) return st;
  }
  if (db) {
    OKL_REQUIRE(g, "okl_conv_wgrad_bf16: the fp32 gradient is needed for db");
    st = okl_channel_sum(ctx, (const float*)g, (float*)db, s.N, s.O, (long long)s.OH * s.OW, OKL_NXC);
    if (st != OKL_OK) return st;
  }
  return OKL_OK;
}

// ================================================================================================ small-C convolution on tcgen05
// First layers (video 3 x 3x3x3 -> K = 81, ...): C * taps + 1 <= 96 is far too small for the TMA implicit GEMM above (its k-blocks
// are 32-channel slices of ONE tap) and the FFMA small-K kernels top out at ~25 % of the fp32 peak.  Here the im2col operand is
// BUILT IN SHARED MEMORY by the CTA's own threads, directly in the canonical UMMA layout (K-major rows of 32 tf32 = 128 bytes,
// 128-byte swizzle: 16-byte chunk index XOR (row & 7), 8-row groups 1024 bytes apart), and consumed by tcgen05.mma:
//   fprop : D[128 pixels, O] = A[128 pixels, 96 taps] . W^T ; A built per tile (each thread: one pixel row, 12 x 16-byte chunks),
//           W (96 x O, zero padded) built once per CTA; two TMEM accumulators, so the epilogue of tile i (TMEM lane = pixel ->
//           every channel's store is a coalesced 128-byte line of the NC(D)HW output, bias + activation fused) overlaps the MMAs
//           of tile i+1.
//   wgrad : D[O, 96] = g^T[O, 32 pixels] . patch[96, 32 pixels]^T accumulated in TMEM over the CTA's share of all 32-pixel
//           k-blocks (split-K over CTAs, fixed-order second-stage reduce); the g rows are 128-byte copies, patch row k is the
//           input shifted by tap k (row K is all ones: its column of D is the bias gradient).
// The padded rows / taps are zero, so they cost tensor-core time (irrelevant) but no memory traffic.
namespace {

constexpr int SC_K = 96;               // padded reduction extent (3 k-blocks of 32 tf32)
constexpr int SC_THREADS = 256;

__device__ __forceinline__ uint32_t sw128_off(int r, int c) {   // byte offset of element (row r, col c < 32) in a swizzled K-major tile
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 2) ^ (r & 7))) << 4) + ((c & 3) << 2));
}

struct ScParams {
  const float* x; const float* W; const float* b; float* y;       // fprop
  const float* gy; float* partial;                                 // wgrad
  okl_conv::ConvGeom g;
  okl_conv::FastDiv dSo;
  int K, Mpix, So, act, ntiles, kb_per_img, kb_total, kb_per_cta;
};

struct ScTaps { int off[SC_K]; int pk[SC_K]; };                    // tap k -> offset from the window origin, packed (td, th, tw)

__device__ __forceinline__ void sc_build_taps(ScTaps* t, const ScParams& p) {
  for (int k = threadIdx.x; k < SC_K; k += blockDim.x) {
    int off = 0, pk = 31 << 20;                   // padding taps: bit 31 is never set in a window mask -> value 0, no load
    if (k < p.K) {
      unsigned u, tw, th, td;
      p.g.dKW.divmod((unsigned)k, u, tw); p.g.dKH.divmod(u, u, th); p.g.dKD.divmod(u, u, td);
      off = (int)(u * p.g.xs[1] + td * p.g.xs[2] + th * p.g.xs[3] + tw * p.g.xs[4]);
      pk = (int)((td << 20) | (th << 10) | tw);
    }
    t->off[k] = off;
    t->pk[k] = pk;
  }
}

// bit t of the result = tap offset t of a window starting at i0 lies inside [0, I)   (k <= 16 taps per dimension)
__device__ __forceinline__ unsigned sc_mask(int i0, int k, int I) {
  const int lo = max(0, -i0), hi = min(k, I - i0);
  return hi > lo ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
}
// value of tap k given the window's three validity masks: ONE code path for interior and border pixels (no divergence)
__device__ __forceinline__ float sc_tap_m(const ScTaps* t, const float* xw, unsigned dm, unsigned hm, unsigned wm, int k) {
  const int pk = t->pk[k];
  const unsigned in = (dm >> (pk >> 20)) & (hm >> ((pk >> 10) & 1023)) & (wm >> (pk & 1023)) & 1u;
  return in ? __ldg(xw + t->off[k]) : 0.f;
}

template <int NB>     // NB = padded output channels (UMMA N): 64 or 128
__global__ void __launch_bounds__(SC_THREADS, NB == 64 ? 3 : 2) sc_fprop_kernel(const ScParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_KB = 128 * 128;                 // one k-block of A: 128 pixel rows x 128 bytes
  constexpr int B_KB = NB * 128;
  uint8_t* sA = smem;                             // [3 k-blocks][128 rows]
  uint8_t* sB = smem + 3 * A_KB;                  // [3 k-blocks][NB rows]
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(sB + 3 * B_KB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 2);             // [NB]
  ScTaps* taps = reinterpret_cast<ScTaps*>(bias_s + NB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<NB>(tmem_slot);
  sc_build_taps(taps, p);
  for (int i = tid; i < NB; i += SC_THREADS) bias_s[i] = (p.b && i < p.g.O) ? __ldg(p.b + i) : 0.f;
  for (int i = tid; i < NB * SC_K; i += SC_THREADS) {          // weights, zero padded to NB x 96
    const int o = i / SC_K, k = i - o * SC_K;
    const float v = (o < p.g.O && k < p.K) ? __ldg(p.W + (size_t)o * p.K + k) : 0.f;
    *reinterpret_cast<float*>(sB + (k >> 5) * B_KB + sw128_off(o, k & 31)) = v;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = make_idesc(0, false, false, 128, NB);

  // This thread builds pixel row r of the 128 x 96 patch tile: chunks half*4 .. half*4+3 of each of the 3 k-blocks = 48 taps, as
  // a ROLLED loop over its 12 sixteen-byte chunks (the fully unrolled, register-prefetching version was 64 KB of code and ran at
  // 32 % issue utilisation with "no instruction" as the top stall); the second resident CTA hides the load latency instead.
  const int r = tid & 127, half = tid >> 7;
  uint32_t parity = 0;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    {
      const int m = tile * 128 + r;
      const bool ok = m < p.Mpix;
      unsigned t = 0, ow = 0, oh = 0, od = 0;
      if (ok) { p.g.dOW.divmod((unsigned)m, t, ow); p.g.dOH.divmod(t, t, oh); p.g.dOD.divmod(t, t, od); }
      const int id0 = (int)od * p.g.S[0] - p.g.P[0], ih0 = (int)oh * p.g.S[1] - p.g.P[1], iw0 = (int)ow * p.g.S[2] - p.g.P[2];
      const float* xw = p.x + (size_t)t * p.g.xs[0] + (long long)id0 * p.g.xs[2] + (long long)ih0 * p.g.xs[3] + (long long)iw0 * p.g.xs[4];
      const unsigned dm = ok ? sc_mask(id0, p.g.Kk[0], p.g.I[0]) : 0u, hm = sc_mask(ih0, p.g.Kk[1], p.g.I[1]), wm = sc_mask(iw0, p.g.Kk[2], p.g.I[2]);
#pragma unroll 1
      for (int q = 0; q < 12; q++) {
        const int kb = q >> 2, c = half * 4 + (q & 3), k0 = kb * 32 + c * 4;
        const float v0 = sc_tap_m(taps, xw, dm, hm, wm, k0), v1 = sc_tap_m(taps, xw, dm, hm, wm, k0 + 1);
        const float v2 = sc_tap_m(taps, xw, dm, hm, wm, k0 + 2), v3 = sc_tap_m(taps, xw, dm, hm, wm, k0 + 3);
        *reinterpret_cast<float4*>(sA + kb * A_KB + sw128_off(r, c * 4)) = make_float4(v0, v1, v2, v3);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < 3; kb++) {
        const uint64_t adesc = make_smem_desc(smem_u32(sA + kb * A_KB), 16, 1024, 2);
        const uint64_t bdesc = make_smem_desc(smem_u32(sB + kb * B_KB), 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < 4; k++) umma<0>(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
      }
      umma_commit(&mma_done[0]);
    }
    // ---- epilogue: TMEM lane = pixel, so every channel's store is one coalesced 128-byte line of the NC(D)HW output
    mbar_wait(&mma_done[0], parity);
    parity ^= 1;
    tc_fence_after();
    {
      const int q = warp & 3;
      const int m = tile * 128 + q * 32 + lane;
      unsigned n = 0, sp = 0;
      if (m < p.Mpix) p.dSo.divmod((unsigned)m, n, sp);
      float* yb = p.y + (size_t)n * p.g.ys[0] + sp;
#pragma unroll
      for (int ch = (warp >> 2); ch < NB / 32; ch += 2) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
        tmem_ld_wait();
        if (m < p.Mpix) {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const int o = ch * 32 + j;
            if (o < p.g.O) yb[(size_t)o * p.g.ys[1]] = act_apply(__uint_as_float(v[j]) + bias_s[o], p.act);
          }
        }
      }
    }
    tc_fence_before();                            // the next tile's MMAs overwrite the accumulator only after the sync below
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<NB>(tmem_base);
}

constexpr int SCW_NST = 2;             // smem stages of the wgrad kernel (58 KB -> three CTAs per SM hide each other's load latency)
constexpr int SCW_MAXA = 4;            // 16-byte chunks of g per thread and k-block (O <= 128: 128 * 8 / 256)
constexpr int SCW_MAXB = SC_K / (SC_THREADS / 32);   // patch values per thread and k-block

__global__ void __launch_bounds__(SC_THREADS, 3) sc_wgrad_kernel(const ScParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = 128 * 128, B_BYTES = SC_K * 128, STAGE = A_BYTES + B_BYTES, NST = SCW_NST;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NST * STAGE);    // stage_free[NST], done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NST + 1);
  ScTaps* taps = reinterpret_cast<ScTaps*>(tmem_slot + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i <= NST; i++) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<128>(tmem_slot);
  sc_build_taps(taps, p);
  for (int i = tid; i < NST * STAGE / 16; i += SC_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = make_idesc(0, false, false, 128, SC_K);

  const int kb_begin = blockIdx.x * p.kb_per_cta, kb_end = min(p.kb_total, kb_begin + p.kb_per_cta);
  const int na = (p.g.O * 8 + SC_THREADS - 1) / SC_THREADS;          // g chunks this thread copies per k-block
  // registers holding the NEXT k-block's global data: the loads are issued one iteration ahead, so their latency overlaps the
  // shared-memory stores, the barrier and the MMA issue of the current one (and the other resident CTAs' work)
  float4 av[SCW_MAXA];
  float bv[SCW_MAXB];
  auto fetch = [&](int kb) {
    const int n = kb / p.kb_per_img, p0 = (kb - n * p.kb_per_img) * 32;
#pragma unroll
    for (int q = 0; q < SCW_MAXA; q++) {
      const int idx = tid + q * SC_THREADS, o = idx >> 3, c = idx & 7;
      av[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < na && o < p.g.O && p0 + c * 4 < p.So)
        av[q] = __ldg(reinterpret_cast<const float4*>(p.gy + (size_t)n * p.g.ys[0] + (size_t)o * p.g.ys[1] + p0 + c * 4));
    }
    const int sp = p0 + lane;
    const bool ok = sp < p.So;
    unsigned t = 0, ow = 0, oh = 0, od = 0;
    if (ok) { p.g.dOW.divmod((unsigned)sp, t, ow); p.g.dOH.divmod(t, od, oh); }
    const int id0 = (int)od * p.g.S[0] - p.g.P[0], ih0 = (int)oh * p.g.S[1] - p.g.P[1], iw0 = (int)ow * p.g.S[2] - p.g.P[2];
    const float* xw = p.x + (size_t)n * p.g.xs[0] + (long long)id0 * p.g.xs[2] + (long long)ih0 * p.g.xs[3] + (long long)iw0 * p.g.xs[4];
    const unsigned dm = ok ? sc_mask(id0, p.g.Kk[0], p.g.I[0]) : 0u, hm = sc_mask(ih0, p.g.Kk[1], p.g.I[1]), wm = sc_mask(iw0, p.g.Kk[2], p.g.I[2]);
#pragma unroll
    for (int q = 0; q < SCW_MAXB; q++) {
      const int r = warp + q * (SC_THREADS / 32);                 // < 96: padding taps read as 0, row K is the ones row
      bv[q] = (r == p.K) ? (ok ? 1.f : 0.f) : sc_tap_m(taps, xw, dm, hm, wm, r);
    }
  };
  if (kb_begin < kb_end) fetch(kb_begin);
  int i = 0;
  for (int kb = kb_begin; kb < kb_end; kb++, i++) {
    const int s = i % NST;
    if (i >= NST) mbar_wait(&bars[s], (uint32_t)((i / NST - 1) & 1));      // the MMAs that read this stage NST k-blocks ago are done
    uint8_t* a = smem + s * STAGE;
    uint8_t* b = a + A_BYTES;
    // A: row o = 32 consecutive pixels of channel o of the output gradient (one 128-byte line), as 16-byte chunks
#pragma unroll
    for (int q = 0; q < SCW_MAXA; q++) {
      const int idx = tid + q * SC_THREADS, o = idx >> 3, c = idx & 7;
      if (q < na && o < p.g.O) *reinterpret_cast<float4*>(a + sw128_off(o, c * 4)) = av[q];
    }
    // B: row k = the input shifted by tap k at the same 32 pixels; row K = ones (bias gradient); rows > K stay zero
#pragma unroll
    for (int q = 0; q < SCW_MAXB; q++) {
      const int r = warp + q * (SC_THREADS / 32);
      if (r <= p.K) *reinterpret_cast<float*>(b + sw128_off(r, lane)) = bv[q];
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (kb + 1 < kb_end) fetch(kb + 1);
    if (tid == 0) {
      tc_fence_after();
      const uint64_t adesc = make_smem_desc(smem_u32(a), 16, 1024, 2);
      const uint64_t bdesc = make_smem_desc(smem_u32(b), 16, 1024, 2);
#pragma unroll
      for (int k = 0; k < 4; k++) umma<0>(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (i | k) ? 1u : 0u);
      umma_commit(&bars[s]);
    }
  }
  if (tid == 0) umma_commit(&bars[NST]);          // everything issued so far has completed when this barrier flips
  mbar_wait(&bars[NST], 0);
  tc_fence_after();
  // D[o, k] -> partial[block][o][k]  (TMEM lane = o; three 32-column chunks)
  {
    const int q = warp & 3, o = q * 32 + lane;
    for (int ch = (warp >> 2); ch < SC_K / 32; ch += 2) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
      tmem_ld_wait();
      if (o < p.g.O && kb_begin < kb_end) {
        float* dst = p.partial + ((size_t)blockIdx.x * p.g.O + o) * SC_K + ch * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem_base);
}

// dW[o, k] = sum_blocks partial[b][o][k], db[o] = sum_blocks partial[b][o][K]: one warp per output, fixed order
__global__ void sc_wgrad_reduce(const float* __restrict__ partial, float* __restrict__ dW, float* __restrict__ db, int O, int K, int nblk) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (idx >= O * (K + 1)) return;
  const int o = idx / (K + 1), k = idx - o * (K + 1);
  float s = 0.f;
  for (int b = lane; b < nblk; b += 32) s += partial[((size_t)b * O + o) * SC_K + k];
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
  if (lane == 0) {
    if (k < K) { if (dW) dW[(size_t)o * K + k] = s; }
    else if (db) db[o] = s;
  }
}

int sc_make_params(const okl_conv_desc* d, ScParams& p) {
  int s = okl_conv::make_geom(d, p.g);
  if (s != OKL_OK) return s;
  p.K = p.g.C * p.g.T;
  p.So = p.g.Oe[0] * p.g.Oe[1] * p.g.Oe[2];
  p.Mpix = p.g.N * p.So;
  p.dSo = okl_conv::FastDiv((unsigned)(p.So > 0 ? p.So : 1));
  p.act = d->act;
  return OKL_OK;
}

}  // namespace

// K = C * taps in [48, 95] (smaller K: the direct kernels are already bound by the activation traffic), O <= 128, plain NC(D)HW
// tensors, enough pixels to fill the chip; the 16-byte loads of the gradient need So % 4 == 0.
bool okl_tc_conv_smallc_supported(okl_ctx* ctx, const okl_conv_desc* d) {
  if (!d || ctx->cc < 100 || d->x_layout != OKL_NCX || d->y_layout != OKL_NCX) return false;
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("OKL_TC_SMALLC"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return false;
  long long K = d->C, So = 1, taps_ok = 1;
  for (int i = 0; i < d->nd; i++) {
    K *= d->k[i];
    if (d->k[i] > 16) taps_ok = 0;
    const long long e = (d->in[i] + 2LL * d->p[i] - d->k[i]) / d->s[i] + 1;
    So *= e;
  }
  return taps_ok && K >= 48 && K + 1 <= SC_K && d->O <= 128 && So % 4 == 0 && (long long)d->N * So >= 128LL * ctx->sm_count &&
         (long long)d->N * So < (1LL << 30);
}

int okl_tc_conv_smallc_fprop(okl_ctx* ctx, const okl_conv_desc* d, const float* x, const float* W, const float* b, float* y) {
  ScParams p;
  memset(&p, 0, sizeof(p));
  int s = sc_make_params(d, p);
  if (s != OKL_OK) return s;
  if (p.Mpix == 0) return OKL_OK;
  p.x = x; p.W = W; p.b = b; p.y = y;
  p.ntiles = (p.Mpix + 127) / 128;
  const int per_sm = p.g.O <= 64 ? 3 : 2;                                               // resident CTAs per SM (smem, registers)
  const int grid = p.ntiles < per_sm * ctx->sm_count ? p.ntiles : per_sm * ctx->sm_count;
  if (p.g.O <= 64) {
    const int smem = 3 * 128 * 128 + 3 * 64 * 128 + 64 + 64 * 4 + (int)sizeof(ScTaps) + 1024;
    static bool set = false;
    if (!set) { OKL_CHECK_CUDA(cudaFuncSetAttribute(sc_fprop_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    sc_fprop_kernel<64><<<grid, SC_THREADS, smem, ctx->stream>>>(p);
  } else {
    const int smem = 3 * 128 * 128 + 3 * 128 * 128 + 64 + 128 * 4 + (int)sizeof(ScTaps) + 1024;
    static bool set = false;
    if (!set) { OKL_CHECK_CUDA(cudaFuncSetAttribute(sc_fprop_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    sc_fprop_kernel<128><<<grid, SC_THREADS, smem, ctx->stream>>>(p);
  }
  OKL_LAUNCHED(ctx, "tc_conv_smallc_fprop");
  return OKL_OK;
}

int okl_tc_conv_smallc_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const float* x, const float* g, float* dW, float* db) {
  ScParams p;
  memset(&p, 0, sizeof(p));
  int s = sc_make_params(d, p);
  if (s != OKL_OK) return s;
  p.x = x; p.gy = g;
  p.kb_per_img = (p.So + 31) / 32;
  p.kb_total = p.g.N * p.kb_per_img;
  int grid = 3 * ctx->sm_count;                       // three resident CTAs per SM
  if (grid > p.kb_total) grid = p.kb_total;
  p.kb_per_cta = (p.kb_total + grid - 1) / grid;
  grid = (p.kb_total + p.kb_per_cta - 1) / p.kb_per_cta;
  s = okl_ensure_workspace(ctx, (size_t)grid * p.g.O * SC_K * sizeof(float));
  if (s != OKL_OK) return s;
  p.partial = (float*)ctx->workspace;
  const int smem = SCW_NST * (128 * 128 + SC_K * 128) + 64 + (int)sizeof(ScTaps) + 1024;
  static bool set = false;
  if (!set) { OKL_CHECK_CUDA(cudaFuncSetAttribute(sc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
  sc_wgrad_kernel<<<grid, SC_THREADS, smem, ctx->stream>>>(p);
  OKL_LAUNCHED(ctx, "tc_conv_smallc_wgrad");
  const int total = p.g.O * (p.K + 1);
  sc_wgrad_reduce<<<(total * 32 + 255) / 256, 256, 0, ctx->stream>>>(p.partial, dW, db, p.g.O, p.K, grid);
  OKL_LAUNCHED(ctx, "tc_conv_smallc_wgrad_reduce");
  return OKL_OK;
}
