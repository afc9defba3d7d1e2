// BF16 implicit-GEMM convolution on tcgen05, second generation (fprop + dgrad; okl.py:790-813, 848-883 / 700-711, 722-740).
//
// What the first generation (okl_tc_gemm.cu MODE 1) left on the table, measured on the CIFAR-shaped layers at batch 1024 (ncu,
// profiles/r2_*): tensor pipe 12-29 %, L2 22-33 %, DRAM 6-33 % - nothing saturated.  The per-tile timeline showed a 128 x 64
// tile taking 3.9 us against 0.6 us of MMA time: (1) the single TMA-producer thread did five integer divisions per k-block,
// (2) the 11 000-instruction epilogue (run-time activation switch, scalar fall-backs) missed the instruction cache, (3) every
// filter tap re-loaded its shifted 128-pixel activation tile (9 x 16 KB per 128 x 64 tile) and every 128-pixel tile its own copy
// of the filter bank.  This kernel:
//   * M = 256 pixels per CTA (two 128-row accumulators sharing every filter tile) - half the filter traffic per flop;
//   * SLAB mode (image rows >= 8 pixels wide, tile inside one image): per (channel chunk, horizontal tap v) ONE TMA box of
//     TH + KH - 1 input rows; the KH vertical taps are the same shared-memory slab at a row offset (a multiple of 1024 bytes, so
//     the 128-byte-swizzle phase is unchanged) - 2.4-2.7 x less activation traffic; TMA's out-of-bounds zero fill IS the padding;
//   * independent A (slab / tap tile) and B (filter tile) rings with their own producer warps; no division inside the k loop;
//   * epilogue specialised at compile time (bias + optional ReLU + optional ReLU-backward mask), bf16 packed in registers, staged
//     per warp in the swizzled layout and written with TMA STORES (cp.async.bulk.tensor ... global.shared::cta) that clip partial
//     tiles; the fp32 copy of the output is optional (BF16-only activations between tensor-core layers);
//   * double-buffered TMEM accumulators (2 tiles x 2 sub-tiles x BN columns) so the epilogue overlaps the next tile's MMAs.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int CT_THREADS = 384;        // warp 0: A producer, 1: MMA issuer, 2: TMEM allocator, 3: B producer, 4-11: epilogue
constexpr int CT_A_STAGE = 40960;      // one (TH + KH - 1)-row slab (<= 320 rows of 128 B) or one 256-pixel tap tile (32 KB)
constexpr int CT_NA = 3;
constexpr int CT_STG = 4096;           // per epilogue warp: 32 rows x 128 B, swizzled like the TMA box it is stored from
constexpr int CT_MAX_OC = 512;

template <int BN>
struct CtCfg {
  static constexpr int B_STAGE = BN * 128;
  static constexpr int NB = BN == 64 ? 6 : (BN == 128 ? 4 : 2);
  static constexpr int ACC = (4 * BN <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = ACC * 2 * BN;            // 256 / 512 / 512
  static constexpr int SMEM = CT_NA * CT_A_STAGE + NB * B_STAGE + 8 * CT_STG + 512 /*barriers*/ + CT_MAX_OC * 4 /*bias*/ + 1024 /*align*/;
};

struct ConvTcParams {
  int tiles_m, tiles_n, tiles_w, tiles_h;
  int TW, TH, TN, tw_sh, th_sh;        // CTA tile: TW x TH pixels of TN images = 256 output pixels (all powers of two)
  int KH, KW, cblocks, C;              // taps; 64-channel chunks of the reduction; channels of the input tensor
  int KD, pd, OD;                      // Conv3D (okl.py:948-985): depth taps, depth padding, output depth - an "image" of the tile decode is
                                       // one (sample, output depth) slice, the depth taps are one more loop around the reduction
  int ph, pw;
  int slab;                            // 1: A = one slab per (chunk, v), vertical taps are row offsets into it; 0: one tile per tap
  int sub_bytes, row_bytes, a_bytes;   // second sub-tile / one tap row inside an A stage; bytes of one A transaction
  int OH, OW, NB, Oc;
  int RH, RN;                          // TMA store box of one epilogue warp (32 pixels): TW x RH x RN
  const float* bias;
  int relu;
  const __nv_bfloat16* mask;           // optional, indexed like the output: out = 0 where mask <= 0 (backward of a preceding ReLU)
  int want_f32, want_b16;
  int gdepth;                          // > 0: grouped stages (slab mode): ring depth shared by the slab ring and the filter-group ring
  int a_stage;                         // bytes between activation stages (grouped mode packs them to make room for a third filter group)
  int diag;                            // probe only (OKL_CT_DIAG): 1 = epilogue does nothing, 2 = no MMAs, 4 = activation operand loaded once per stage
  unsigned long long* dbg;
};

template <int BN, int PAIR>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY16,
               const __grid_constant__ CUtensorMap tmY32, const ConvTcParams p) {
  using Cfg = CtCfg<BN>;
  constexpr int CS = PAIR ? 2 : 1;
  // filter-tile ring: in pair mode a CTA holds only half of every tile's rows, so the same shared memory gives twice the stages.
  // (Probes, scripts/exp/ct_diag.py + commit_bench.cu: with no MMAs, no epilogue and no loads at all the 64 -> 64 layer still takes 57 of
  // its 94 us - the single issuing thread pays ~200-300 cycles per tcgen05.commit and ~85 per mbarrier wait, 12 of each per tile; the
  // ring depth is not what limits it.)
  constexpr int NBK = PAIR ? 2 * Cfg::NB : Cfg::NB;
  constexpr int BST = PAIR ? Cfg::B_STAGE / 2 : Cfg::B_STAGE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + CT_NA * p.a_stage;
  uint8_t* sStg = sA + CT_NA * CT_A_STAGE + Cfg::NB * Cfg::B_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + 8 * CT_STG);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + CT_NA;
  uint64_t* b_full = a_empty + CT_NA;
  uint64_t* b_empty = b_full + NBK;
  uint64_t* t_full = b_empty + NBK;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  volatile uint32_t* loop_par = reinterpret_cast<volatile uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 384);   // up to 34 barriers + the TMEM slot before it
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 0] = gtime();

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY16);
    tma_prefetch_desc(&tmY32);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < CT_NA; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NBK; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    // pair mode: the even CTA's accumulator-free barrier collects the epilogue warps of BOTH CTAs
    for (int a = 0; a < 2; a++) { mbar_init(&t_full[a], 1); mbar_init(&t_empty[a], PAIR ? 16 : 8); }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_cg2<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  for (int i = threadIdx.x; i < CT_MAX_OC; i += CT_THREADS) bias_s[i] = (p.bias && i < p.Oc) ? __ldg(p.bias + i) : 0.f;
  if (threadIdx.x == 0) {
    loop_par[0] = p.slab ? 1u : 0u;
    loop_par[1] = (uint32_t)p.sub_bytes >> 4;
    loop_par[2] = p.slab ? (uint32_t)p.row_bytes >> 4 : 0u;
    loop_par[3] = (uint32_t)p.KH;
    loop_par[4] = (uint32_t)(p.KD * p.cblocks * p.KW);
    loop_par[5] = p.dbg ? 1u : 0u;
    loop_par[6] = (uint32_t)p.diag;
    loop_par[7] = (uint32_t)p.gdepth;
    loop_par[8] = (uint32_t)p.a_stage >> 4;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  okl_pdl_sync();
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  if (PAIR) cluster_sync_all();                             // both CTAs' barriers and TMEM exist before any cross-CTA signal

  // work units: (output-channel tile, group of CS consecutive pixel tiles); cluster c takes units c, c + #clusters, ...; CTA rank r
  // of the pair takes pixel tile group * CS + r
  const int total = (p.tiles_m / CS) * p.tiles_n;
  const int unit0 = blockIdx.x / CS, unit_step = gridDim.x / CS;

  if (warp == 0) {
    // ===================================================================== A producer: activation slabs / tap tiles (every CTA its own)
    if (elect_one()) {
      int st = 0;
      uint32_t ph_ = 0;
      int nloads = 0;
      const int adepth = p.gdepth > 0 ? p.gdepth : CT_NA;
      auto load = [&](int c0, int c1, int c2, int c3, int c4) {
        mbar_wait(&a_empty[st], ph_ ^ 1);
        if ((p.diag & 4) && nloads++ >= CT_NA) {
          if (!PAIR || crank == 0) mbar_arrive(&a_full[st]);
          if (++st == adepth) { st = 0; ph_ ^= 1; }
          return;
        }
        if (PAIR) {
          if (crank == 0) mbar_expect_tx(&a_full[st], 2 * p.a_bytes);       // both CTAs' boxes are counted on the even CTA's barrier
          if (p.OD > 0) tma_load_5d_pair(&tmA, &a_full[st], sA + st * p.a_stage, c0, c1, c2, c3, c4);
          else tma_load_4d_pair(&tmA, &a_full[st], sA + st * p.a_stage, c0, c1, c2, c3);
        } else {
          mbar_expect_tx(&a_full[st], p.a_bytes);
          if (p.OD > 0) tma_load_5d(&tmA, &a_full[st], sA + st * p.a_stage, c0, c1, c2, c3, c4);
          else tma_load_4d(&tmA, &a_full[st], sA + st * p.a_stage, c0, c1, c2, c3);
        }
        if (++st == adepth) { st = 0; ph_ ^= 1; }
      };
      for (int w = unit0; w < total; w += unit_step) {
        const int tm = (w / p.tiles_n) * CS + (int)crank;
        const int wt = tm % p.tiles_w, r2 = tm / p.tiles_w;
        const int ht = r2 % p.tiles_h, ng = r2 / p.tiles_h;
        const int w0 = wt * p.TW - p.pw, h0 = ht * p.TH - p.ph, n0 = ng * p.TN;
        // 3-D: image group -> (sample, first output depth); the 5-D box takes TN consecutive depths of that sample (TN divides OD)
        const int smp = p.OD > 0 ? n0 / p.OD : 0, d0 = p.OD > 0 ? n0 - smp * p.OD - p.pd : 0;
        for (int t = 0; t < p.KD; t++) {
          for (int cb = 0; cb < p.cblocks; cb++) {
            for (int v = 0; v < p.KW; v++) {
              if (p.slab) load(cb * 64, w0 + v, h0, p.OD > 0 ? d0 + t : n0, smp);
              else
                for (int u = 0; u < p.KH; u++) load(cb * 64, w0 + v, h0 + u, p.OD > 0 ? d0 + t : n0, smp);
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================================================================== B producer: one filter tile per (chunk, tap); in pair mode
    // every CTA loads its half of the tile's N rows
    if (elect_one()) {
      int st = 0;
      uint32_t ph_ = 0;
      for (int w = unit0; w < total; w += unit_step) {
        const int tn = w % p.tiles_n;
        const int n_off = tn * BN + (PAIR ? (int)crank * (BN / 2) : 0);
        for (int t = 0; t < p.KD; t++)
        for (int cb = 0; cb < p.cblocks; cb++) {
          for (int v = 0; v < p.KW; v++) {
            int kcol = (t * p.KH * p.KW + v) * p.C + cb * 64;     // tap (t, u, v) -> column ((t * KH + u) * KW + v) * C + cb * 64 of the K-major bank
            if (p.gdepth > 0) {
              // grouped stages: the KH filter tiles of this (chunk, v) land in ONE group slot, counted on one barrier; the slot is
              // released together with the activation slab of the same iteration (the slab ring's empty barrier)
              mbar_wait(&a_empty[st], ph_ ^ 1);
              if (!PAIR || crank == 0) mbar_expect_tx(&b_full[st], (uint32_t)(p.KH * Cfg::B_STAGE));
              uint8_t* dst = sB + st * (p.KH * BST);
              for (int u = 0; u < p.KH; u++, kcol += p.KW * p.C) {
                if (PAIR) tma_load_2d_pair(&tmB, &b_full[st], dst + u * BST, kcol, n_off);
                else tma_load_2d(&tmB, &b_full[st], dst + u * BST, kcol, n_off);
              }
              if (++st == p.gdepth) { st = 0; ph_ ^= 1; }
              continue;
            }
            for (int u = 0; u < p.KH; u++, kcol += p.KW * p.C) {
              if (p.diag & 8) continue;                    // probe: no filter-tile ring at all (neither loads nor its barriers)
              mbar_wait(&b_empty[st], ph_ ^ 1);
              if (PAIR) {
                if (crank == 0) mbar_expect_tx(&b_full[st], Cfg::B_STAGE);
                tma_load_2d_pair(&tmB, &b_full[st], sB + st * BST, kcol, n_off);
              } else {
                mbar_expect_tx(&b_full[st], Cfg::B_STAGE);
                tma_load_2d(&tmB, &b_full[st], sB + st * BST, kcol, n_off);
              }
              if (++st == NBK) { st = 0; ph_ ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (pair mode: the even CTA only, for both)
    // A lean single-thread loop: per tap one barrier wait, one multiply-add per descriptor low word, eight MMAs, one commit.  (The first
    // version rebuilt three 64-bit descriptors per tap, re-read run-time flags from constant memory and fenced after every wait - ~100
    // dependent scalar instructions per tap against 256 cycles of tensor work at N = 64: issuing every MMA twice cost only 13 %.)
    if (crank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc(1, false, false, PAIR ? 256 : 128, BN);
      constexpr uint32_t HI = TC_DESC_HI_SW128;
      constexpr uint32_t B16 = BST >> 4;
      const uint32_t A16 = loop_par[8];
      const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);      // LBO = 16 B (unused by K-major swizzled operands)
      const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t af0 = smem_u32(a_full), ae0 = smem_u32(a_empty), bf0 = smem_u32(b_full), be0 = smem_u32(b_empty);
      const uint32_t tf0 = smem_u32(t_full), te0 = smem_u32(t_empty);
      // run-time parameters of the loop come from shared memory (written in the prologue): read from the kernel parameters, ptxas
      // re-loads them from constant memory in every iteration, and each of those loads stalls the chain
      const bool slab = loop_par[0] != 0;
      const uint32_t sub16 = loop_par[1], row16 = loop_par[2];
      const int KH = (int)loop_par[3], nv = (int)loop_par[4];
      uint32_t sa = 0, sb = 0, acc = 0;
      uint32_t pa = 0, pb = 0, pacc = 0;
      bool want_ts = loop_par[5] != 0;                     // debug timeline: when the first activation slab landed
      const bool nomma = (loop_par[6] & 2u) != 0, nobring = (loop_par[6] & 8u) != 0;
      const uint32_t gdepth = loop_par[7], g16 = (uint32_t)KH * B16;
      auto commit = [&](uint32_t bar) {
        if (PAIR) umma_commit_pair_a(bar);
        else umma_commit_a(bar);
      };
      for (int w = unit0; w < total; w += unit_step) {
        mbar_wait_a(te0 + acc * 8, pacc ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * (uint32_t)(2 * BN);
        uint32_t accum = 0;
        if (gdepth) {
          // grouped stages (slab mode): per (chunk, v) iteration two waits (slab, filter group), KH x 8 MMAs and ONE commit that frees
          // slab and filter group together - the issuing thread pays ~200-300 cycles per tcgen05.commit and ~85 per barrier wait
          // (scripts/exp/commit_bench.cu), which the per-tap protocol below spends 12 times per 3 x 3 tile
#pragma unroll 1
          for (int cv = 0; cv < nv; cv++) {
            const uint32_t a_lo = a_lo0 + sa * A16, bg = b_lo0 + sa * g16;
            mbar_wait_a(af0 + sa * 8, pa);
            mbar_wait_a(bf0 + sa * 8, pa);
            if (want_ts) {
              p.dbg[blockIdx.x * 8 + 1] = gtime();
              want_ts = false;
            }
#pragma unroll 1
            for (int u = 0; u < KH; u++) {
              const uint32_t al = a_lo + (uint32_t)u * row16, b_lo = bg + (uint32_t)u * B16;
              if (!nomma) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  umma_bf16_words<PAIR>(d0, al + (uint32_t)(k * 2), b_lo + (uint32_t)(k * 2), HI, idesc, k ? 1u : accum);
                  umma_bf16_words<PAIR>(d0 + BN, al + sub16 + (uint32_t)(k * 2), b_lo + (uint32_t)(k * 2), HI, idesc, k ? 1u : accum);
                }
              }
              accum = 1;
            }
            commit(ae0 + sa * 8);
            if (++sa == gdepth) { sa = 0; pa ^= 1; }
          }
        } else
#pragma unroll 1
        for (int cv = 0; cv < nv; cv++) {                  // (channel chunk, horizontal tap)
          uint32_t a_lo = a_lo0 + sa * A16;
          if (slab) {
            mbar_wait_a(af0 + sa * 8, pa);
            if (want_ts) {
              p.dbg[blockIdx.x * 8 + 1] = gtime();
              want_ts = false;
            }
          }
#pragma unroll 1
          for (int u = 0; u < KH; u++) {
            const uint32_t b_lo = b_lo0 + sb * B16;
            if (!slab) {
              a_lo = a_lo0 + sa * A16;
              mbar_wait_a(af0 + sa * 8, pa);
            }
            if (!nobring) mbar_wait_a(bf0 + sb * 8, pb);
            const uint32_t al = a_lo + (uint32_t)u * row16;
            if (!nomma) {
#pragma unroll
              for (int k = 0; k < 4; k++) {
                umma_bf16_words<PAIR>(d0, al + (uint32_t)(k * 2), b_lo + (uint32_t)(k * 2), HI, idesc, k ? 1u : accum);
                umma_bf16_words<PAIR>(d0 + BN, al + sub16 + (uint32_t)(k * 2), b_lo + (uint32_t)(k * 2), HI, idesc, k ? 1u : accum);
              }
            }
            accum = 1;
            if (!nobring) commit(be0 + sb * 8);
            if (++sb == (uint32_t)NBK) { sb = 0; pb ^= 1; }
            if (!slab) {
              commit(ae0 + sa * 8);
              if (++sa == (uint32_t)CT_NA) { sa = 0; pa ^= 1; }
            }
          }
          if (slab) {
            commit(ae0 + sa * 8);
            if (++sa == (uint32_t)CT_NA) { sa = 0; pa ^= 1; }
          }
        }
        commit(tf0 + acc * 8);
        if (Cfg::ACC == 2) { acc ^= 1; if (acc == 0) pacc ^= 1; }
        else pacc ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue: warp ew serves sub-tile ew / 4, TMEM lanes 32 * (ew % 4) ..
    const int ew = warp - 4, s = ew >> 2, q = ew & 3;
    uint8_t* stg = sStg + ew * CT_STG;
    const int pbase = s * 128 + q * 32;                   // first pixel (row of the CTA tile) of this warp
    const int r = pbase + lane;
    int acc = 0, dbg_tile = 0;
    uint32_t pacc = 0;
    const uint32_t sw = (uint32_t)(lane & 7);
    for (int w = unit0; w < total; w += unit_step) {
      const int tn = w % p.tiles_n;
      const int tm = (w / p.tiles_n) * CS + (int)crank;
      const int wt = tm % p.tiles_w, r2 = tm / p.tiles_w;
      const int ht = r2 % p.tiles_h, ng = r2 / p.tiles_h;
      const int w0 = wt * p.TW, h0 = ht * p.TH, n0 = ng * p.TN;
      const int ch0 = tn * BN;
      const int bh = h0 + ((pbase >> p.tw_sh) & (p.TH - 1)), bn = n0 + (pbase >> (p.tw_sh + p.th_sh));
      const __nv_bfloat16* mrow = nullptr;
      if (p.mask) {
        const int rw = w0 + (r & (p.TW - 1)), rh = h0 + ((r >> p.tw_sh) & (p.TH - 1)), rn = n0 + (r >> (p.tw_sh + p.th_sh));
        if (rw < p.OW && rh < p.OH && rn < p.NB) mrow = p.mask + (((size_t)rn * p.OH + rh) * p.OW + rw) * (size_t)p.Oc + ch0;
      }
      // the ReLU mask of the first 64 channels is requested BEFORE the wait for the accumulator: its HBM latency hides behind the MMAs
      uint4 mk[8];
      const int crem = p.Oc - ch0;                         // output channels left from this tile's first one (partial last tile)
      if (p.mask) {
#pragma unroll
        for (int j = 0; j < 8; j++) mk[j] = (mrow && j * 8 < crem) ? __ldg(reinterpret_cast<const uint4*>(mrow) + j) : make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(&t_full[acc], pacc);
      tc_fence_after();
      if (p.dbg && threadIdx.x == 128 && blockIdx.x == 0 && dbg_tile < 24) p.dbg[148 * 8 + 2 * dbg_tile] = gtime();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * BN + s * BN);
#pragma unroll 1
      for (int ch = 0; ch < ((p.diag & 1) ? 0 : BN / 64); ch++) {
        if (p.mask && ch > 0) {
#pragma unroll
          for (int j = 0; j < 8; j++)
            mk[j] = (mrow && ch * 64 + j * 8 < crem) ? __ldg(reinterpret_cast<const uint4*>(mrow + ch * 64) + j) : make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr + (uint32_t)(ch * 64), v0);
        tmem_ld32(taddr + (uint32_t)(ch * 64 + 32), v1);
        tmem_ld_wait();
        const float* bz = bias_s + ch0 + ch * 64;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          float a = __uint_as_float(v0[j]) + bz[j], b = __uint_as_float(v1[j]) + bz[32 + j];
          if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
          v0[j] = __float_as_uint(a);
          v1[j] = __float_as_uint(b);
        }
        if (p.mask) {
          // bf16 sign/zero test on the raw 16-bit patterns: value > 0  <=>  sign bit clear and not +0
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const uint32_t wds[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const uint32_t lo = wds[e] & 0xFFFFu, hi = wds[e] >> 16;
              const int c = j * 8 + e * 2;                                   // channel inside the 64-channel chunk
              const bool klo = (lo & 0x8000u) == 0 && (lo & 0x7FFFu) != 0, khi = (hi & 0x8000u) == 0 && (hi & 0x7FFFu) != 0;
              if (c < 32) { if (!klo) v0[c] = 0u; if (!khi) v0[c + 1] = 0u; }
              else { if (!klo) v1[c - 32] = 0u; if (!khi) v1[c - 31] = 0u; }
            }
          }
        }
        if (p.want_f32) {
#pragma unroll
          for (int half = 0; half < 2; half++) {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
            const uint32_t* vv = half ? v1 : v0;
#pragma unroll
            for (int j = 0; j < 8; j++)
              *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ sw) << 4)) = make_uint4(vv[4 * j], vv[4 * j + 1], vv[4 * j + 2], vv[4 * j + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmY32, stg, ch0 + ch * 64 + half * 32, w0, bh, bn);
              tma_store_commit();
            }
          }
        }
        if (p.want_b16) {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const uint32_t* vv = j < 4 ? v0 : v1;
            const int b = (j & 3) * 8;
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(vv[b + 2 * e]), __uint_as_float(vv[b + 2 * e + 1]));
              pk[e] = *reinterpret_cast<uint32_t*>(&t);
            }
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmY16, stg, ch0 + ch * 64, w0, bh, bn);
            tma_store_commit();
          }
        }
      }
      if (p.dbg && threadIdx.x == 128 && blockIdx.x == 0 && dbg_tile < 24) p.dbg[148 * 8 + 2 * dbg_tile + 1] = gtime();
      dbg_tile++;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&t_empty[acc]), 0u));   // the issuing (even) CTA waits for both epilogues
        else mbar_arrive(&t_empty[acc]);
      }
      if (Cfg::ACC == 2) { if (++acc == 2) { acc = 0; pacc ^= 1; } }
      else pacc ^= 1;
    }
    if (lane == 0) tma_store_wait<0>();                    // every store of this warp has been written before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                             // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_cg2<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 3] = gtime();
}

// ------------------------------------------------------------------------------------------------ filter banks
// w2[o][tap][c] = bf16(W[o][c][tap])           K-major bank of the forward convolution
// wf[c][tap][o] = bf16(W[o][c][T - 1 - tap])   flipped + transposed bank: dgrad as a convolution of the output gradient
__global__ void conv_filter_banks_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ w2, __nv_bfloat16* __restrict__ wf, int O, int C,
                                         int T) {
  const size_t total = (size_t)O * C * T;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    {
      const int c = (int)(i % C);
      const size_t r = i / C;
      const int t = (int)(r % T), o = (int)(r / T);
      if (w2) w2[i] = __float2bfloat16(W[((size_t)o * C + c) * T + t]);
    }
    {
      const int o = (int)(i % O);
      const size_t r = i / O;
      const int t = (int)(r % T), c = (int)(r / T);
      if (wf) wf[i] = __float2bfloat16(W[((size_t)o * C + c) * T + (T - 1 - t)]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*CtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

CtEncodeFn ct_encode_fn() {
  static CtEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<CtEncodeFn>(q);
  }
  return fn;
}

// 4-D map over a channels-last tensor (N, H, W, C) of `es`-byte elements: dims {C, W, H, N}, 128-byte swizzle
int ct_map_nhwc(CUtensorMap* map, const void* base, int es, int C, int W, int H, int N, int box_c, int box_w, int box_h, int box_n) {
  CtEncodeFn fn = ct_encode_fn();
  if (!fn) {
    okl_set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return OKL_ERR_CUDA;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstride[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    okl_set_error("cuTensorMapEncodeTiled(nhwc) failed with CUresult %d (C=%d W=%d H=%d N=%d box=%d,%d,%d,%d es=%d)", (int)r, C, W, H, N, box_c,
                  box_w, box_h, box_n, es);
    return OKL_ERR_CUDA;
  }
  return OKL_OK;
}

// 5-D map over a channels-last bf16 tensor (N, D, H, W, C): dims {C, W, H, D, N}; one box = box_d depth slices of one sample
int ct_map_ndhwc(CUtensorMap* map, const void* base, int C, int W, int H, int D, int N, int box_c, int box_w, int box_h, int box_d) {
  CtEncodeFn fn = ct_encode_fn();
  if (!fn) {
    okl_set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return OKL_ERR_CUDA;
  }
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gstride[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_d, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    okl_set_error("cuTensorMapEncodeTiled(ndhwc) failed with CUresult %d (C=%d W=%d H=%d D=%d N=%d box=%d,%d,%d,%d)", (int)r, C, W, H, D, N, box_c, box_w,
                  box_h, box_d);
    return OKL_ERR_CUDA;
  }
  return OKL_OK;
}

// 2-D map over a row-major bf16 matrix [outer][inner]
int ct_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer) {
  CtEncodeFn fn = ct_encode_fn();
  if (!fn) {
    okl_set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return OKL_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    okl_set_error("cuTensorMapEncodeTiled(2d) failed with CUresult %d (inner=%llu outer=%llu box=%ux%u)", (int)r, (unsigned long long)inner,
                  (unsigned long long)outer, box_inner, box_outer);
    return OKL_ERR_CUDA;
  }
  return OKL_OK;
}

int ct_pow2_ceil(int v) { int q = 1; while (q < v) q <<= 1; return q; }
int ct_ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }
int ct_env(const char* name, int dflt) { const char* s = getenv(name); return s && *s ? atoi(s) : dflt; }

template <int BN, int PAIR>
int ct_launch(okl_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ty16, const CUtensorMap& ty32, const ConvTcParams& p) {
  using Cfg = CtCfg<BN>;
  constexpr int CS = PAIR ? 2 : 1;
  auto kern = conv_tc_kernel<BN, PAIR>;
  static bool attr_set = false;
  if (!attr_set) {
    OKL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const int units = (p.tiles_m / CS) * p.tiles_n;
  int clusters = ctx->sm_count / CS;
  if (clusters > units) clusters = units;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * CS); cfg.blockDim = dim3(CT_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  memset(at, 0, sizeof(at));
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = PAIR ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, ta, tb, ty16, ty32, p);
  OKL_LAUNCHED(ctx, "conv_tc_kernel");
  return OKL_OK;
}

template <int BN>
int ct_launch_cs(okl_ctx* ctx, int pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ty16, const CUtensorMap& ty32,
                 const ConvTcParams& p) {
  if (pair) return ct_launch<BN, 1>(ctx, ta, tb, ty16, ty32, p);
  return ct_launch<BN, 0>(ctx, ta, tb, ty16, ty32, p);
}

struct CtShape { int N, C, O, H, W, KH, KW, PH, PW, OH, OW; int D, KD, PD, OD; };   // D = KD = OD = 1, PD = 0 for 1-D / 2-D layers
bool gs_feasible(okl_ctx* ctx, const okl_conv_desc* d);

void ct_shape(const okl_conv_desc* d, CtShape& s) {
  s.N = d->N; s.C = d->C; s.O = d->O;
  s.D = s.KD = s.OD = 1; s.PD = 0;
  if (d->nd == 1) { s.H = 1; s.W = d->in[0]; s.KH = 1; s.KW = d->k[0]; s.PH = 0; s.PW = d->p[0]; }
  else if (d->nd == 2) { s.H = d->in[0]; s.W = d->in[1]; s.KH = d->k[0]; s.KW = d->k[1]; s.PH = d->p[0]; s.PW = d->p[1]; }
  else {
    s.D = d->in[0]; s.H = d->in[1]; s.W = d->in[2]; s.KD = d->k[0]; s.KH = d->k[1]; s.KW = d->k[2]; s.PD = d->p[0]; s.PH = d->p[1]; s.PW = d->p[2];
    s.OD = s.D + 2 * s.PD - s.KD + 1;
  }
  s.OH = s.H + 2 * s.PH - s.KH + 1;
  s.OW = s.W + 2 * s.PW - s.KW + 1;
}

// y[N, OH, OW, Oc] = act(conv(x16[N, H, W, Cc], bank[Oc][T * Cc]) + bias) (* mask > 0), stride 1, padding (ph, pw), channels-last
// images per 256-pixel tile of an OH x OW output (1 when the tile lies inside one image)
int ct_tile_images(int OH, int OW) {
  int tw = ct_pow2_ceil(OW); if (tw > 32) tw = 32;
  int th = ct_pow2_ceil(OH); if (th > 256 / tw) th = 256 / tw;
  return 256 / (tw * th);
}

// vol3 != NULL: Conv3D - {D, KD, pd, OD}: input depth, depth taps, depth padding, output depth; N then counts SAMPLES and every
// (sample, output depth) slice is an image of the tile decode
int conv_tc_run(okl_ctx* ctx, const void* x16, int N, int H, int W, int Cc, const void* bank, int Oc, int KH, int KW, int ph, int pw, int OH, int OW,
                const float* bias, int relu, float* y, void* y16, const void* mask16, const int* vol3 = nullptr) {
  OKL_REQUIRE(Cc % 8 == 0 && Oc % 8 == 0 && Oc <= CT_MAX_OC, "conv_tc: channel counts must be multiples of 8 (C=%d O=%d)", Cc, Oc);
  OKL_REQUIRE(y || y16, "conv_tc: no output requested");
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  const int D3 = vol3 ? vol3[0] : 1, KD = vol3 ? vol3[1] : 1, pd3 = vol3 ? vol3[2] : 0, OD = vol3 ? vol3[3] : 1;
  const int samples = N;
  p.KD = KD; p.pd = pd3; p.OD = vol3 ? OD : 0;
  if (vol3) {
    OKL_REQUIRE(OD % ct_tile_images(OH, OW) == 0, "conv_tc (3-D): the images of a tile must be depth slices of one sample (OD=%d)", OD);
    N = samples * OD;                                        // images of the tile decode and of the (dense) output
  }
  p.TW = ct_pow2_ceil(OW); if (p.TW > 32) p.TW = 32;
  p.TH = ct_pow2_ceil(OH); if (p.TH > 256 / p.TW) p.TH = 256 / p.TW;
  p.TN = 256 / (p.TW * p.TH);
  p.tw_sh = ct_ilog2(p.TW); p.th_sh = ct_ilog2(p.TH);
  p.tiles_w = (OW + p.TW - 1) / p.TW; p.tiles_h = (OH + p.TH - 1) / p.TH;
  const int tiles_ng = (N + p.TN - 1) / p.TN;
  p.tiles_m = p.tiles_w * p.tiles_h * tiles_ng;
  // N tile: as wide as the channel count allows (the wider the MMA, the less shared-memory bandwidth per flop: measured 1 239 vs
  // 938 TFLOP/s for BN = 256 vs 128 on the 256 -> 256 layer) as long as the grid still fills the chip
  int bn = ct_env("OKL_CT_BN", 0);
  if (bn != 64 && bn != 128 && bn != 256) {
    bn = Oc % 256 == 0 ? 256 : (Oc % 128 == 0 ? 128 : 64);
    while (bn > 64 && (long long)p.tiles_m * (Oc / bn) * 4 < 3LL * ctx->sm_count) bn >>= 1;
  }
  if (Oc % bn) bn = 64;
  p.tiles_n = (Oc + bn - 1) / bn;
  p.KH = KH; p.KW = KW; p.cblocks = (Cc + 63) / 64; p.C = Cc; p.ph = ph; p.pw = pw;
  p.row_bytes = p.TW * 128;
  p.slab = (p.TN == 1 && KH > 1 && p.TW >= 8 && (p.TH + KH - 1) * p.TW * 128 <= CT_A_STAGE && ct_env("OKL_CT_SLAB", 1)) ? 1 : 0;
  p.sub_bytes = p.slab ? (p.TH / 2) * p.row_bytes : 128 * 128;
  p.a_bytes = p.slab ? (p.TH + KH - 1) * p.row_bytes : 256 * 128;
  p.OH = OH; p.OW = OW; p.NB = N; p.Oc = Oc;
  p.RH = p.TH < 32 / p.TW ? p.TH : 32 / p.TW; if (p.RH < 1) p.RH = 1;
  p.RN = 32 / (p.TW * p.RH);
  p.bias = bias; p.relu = relu; p.mask = (const __nv_bfloat16*)mask16;
  p.want_f32 = y ? 1 : 0; p.want_b16 = y16 ? 1 : 0;
  p.diag = ct_env("OKL_CT_DIAG", 0);
  p.dbg = (unsigned long long*)ctx->tc_dbg;
  // CTA pairs (cta_group::2): two CTAs share every filter tile (each holds half of its rows); needs an even number of pixel tiles
  // measured (batch 1024, us, pair vs single): 64->64 fprop 98 vs 120, dgrad 123 vs 160; 128->128 58 vs 70 (1 322 TFLOP/s), 68 vs 82;
  // at BN = 256 the operand fetch is no longer the limit (61.7 single vs 64.5 pair) - those tiles stay single
  int pair = ct_env("OKL_CT_PAIR", bn <= 128 ? 1 : 0) ? 1 : 0;
  if (p.tiles_m % 2 != 0 || (p.tiles_m / 2) * p.tiles_n < ctx->sm_count / 4) pair = 0;
  // grouped stages: slab mode, and at least two filter groups (KH tiles of this CTA's rows) fit the filter ring
  // (measured, batch 1024: 64 -> 64 @ 32 x 32 96 -> 88 us; with only two groups the 128 -> 128 layer got slower, 55 -> 60 us, so
  // three are required - the activation stages are packed to their real size to make room for the third)
  p.gdepth = 0;
  p.a_stage = CT_A_STAGE;
  if (p.slab && ct_env("OKL_CT_GROUP", 1) && !(p.diag & 8)) {
    const int bst = bn * 128 / (pair ? 2 : 1);
    const int a_stage = (p.a_bytes + 1023) & ~1023;
    const int ring = (bn == 64 ? 6 : (bn == 128 ? 4 : 2)) * bn * 128 + CT_NA * (CT_A_STAGE - a_stage);
    if (ring / (KH * bst) >= CT_NA) { p.gdepth = CT_NA; p.a_stage = a_stage; }
  }
  CUtensorMap ta, tb, ty16, ty32;
  int s = vol3 ? ct_map_ndhwc(&ta, x16, Cc, W, H, D3, samples, 64, p.TW, p.slab ? p.TH + KH - 1 : p.TH, p.TN)
               : ct_map_nhwc(&ta, x16, 2, Cc, W, H, N, 64, p.TW, p.slab ? p.TH + KH - 1 : p.TH, p.TN);
  if (s != OKL_OK) return s;
  s = ct_map_2d(&tb, bank, (uint64_t)KD * KH * KW * Cc, (uint64_t)Oc, 64, (uint32_t)(pair ? bn / 2 : bn));
  if (s != OKL_OK) return s;
  // an absent output still needs a valid descriptor object (never dereferenced by the kernel): alias the other one
  s = ct_map_nhwc(&ty16, y16 ? y16 : (const void*)y, y16 ? 2 : 4, Oc, OW, OH, N, y16 ? 64 : 32, p.TW, p.RH, p.RN);
  if (s != OKL_OK) return s;
  s = ct_map_nhwc(&ty32, y ? (const void*)y : y16, y ? 4 : 2, Oc, OW, OH, N, y ? 32 : 64, p.TW, p.RH, p.RN);
  if (s != OKL_OK) return s;
  switch (bn) {
    case 64: return ct_launch_cs<64>(ctx, pair, ta, tb, ty16, ty32, p);
    case 128: return ct_launch_cs<128>(ctx, pair, ta, tb, ty16, ty32, p);
    default: return ct_launch_cs<256>(ctx, pair, ta, tb, ty16, ty32, p);
  }
}

bool ct_supported(okl_ctx* ctx, const okl_conv_desc* d) {
  if (ctx->cc < 100 || ct_env("OKL_CONV_TC2", 1) == 0) return false;
  if (d->nd > 3 || d->x_layout != OKL_NXC || d->y_layout != OKL_NXC) return false;
  for (int i = 0; i < d->nd; i++)
    if (d->s[i] != 1) return false;
  if (d->C % 8 || d->O % 8 || d->C > CT_MAX_OC || d->O > CT_MAX_OC || d->N < 1) return false;
  CtShape s;
  ct_shape(d, s);
  if (s.OH < 1 || s.OW < 1 || s.KH > 7 || s.KW > 7) return false;
  if (s.PH > s.KH - 1 || s.PW > s.KW - 1) return false;
  if (d->nd == 3) {
    // Conv3D: a tile's images are depth slices of ONE sample, in the forward (output depth) and in the input gradient (input depth);
    // the weight gradient comes from the gradient-slab kernel, one depth tap per CTA group
    if (ct_env("OKL_CT_3D", 1) == 0 || s.OD < 1 || s.KD > 7 || s.PD > s.KD - 1) return false;
    if (s.OD % ct_tile_images(s.OH, s.OW) || s.D % ct_tile_images(s.H, s.W)) return false;
    if (!gs_feasible(ctx, d)) return false;
  }
  // Channel counts that are multiples of 8 but not of 64 (16-byte rows for TMA): the 64-channel boxes run past the tensor's channel
  // extent and TMA's zero fill pads the reduction; partial output-channel tiles are clipped by the TMA stores.  Those shapes need the
  // gradient-slab weight-gradient kernel (the other wgrad kernels assume whole 64-channel chunks), so they are taken only with it.
  if ((d->C % 64 || d->O % 64) && (ct_env("OKL_CT_ANYC", 1) == 0 || !gs_feasible(ctx, d))) return false;
  return true;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI (include/okl.h)
extern "C" int okl_conv2_supported(okl_ctx* ctx, const okl_conv_desc* d) {
  if (!ctx || !d) return 0;
  return ct_supported(ctx, d) ? 1 : 0;
}

extern "C" int okl_conv2_filter_banks(okl_ctx* ctx, const okl_conv_desc* d, const void* W, void* w2_16, void* wf_16) {
  OKL_REQUIRE(ctx && d && W && (w2_16 || wf_16), "NULL argument");
  long long T = 1;
  for (int i = 0; i < d->nd; i++) T *= d->k[i];
  const size_t total = (size_t)d->O * d->C * T;
  conv_filter_banks_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)W, (__nv_bfloat16*)w2_16, (__nv_bfloat16*)wf_16, d->O,
                                                                                 d->C, (int)T);
  OKL_LAUNCHED(ctx, "conv_filter_banks");
  return OKL_OK;
}

extern "C" int okl_conv2_fprop(okl_ctx* ctx, const okl_conv_desc* d, const void* x16, const void* w2_16, const void* bias, void* y, void* y16) {
  OKL_REQUIRE(ctx && d && x16 && w2_16 && (y || y16), "NULL argument");
  OKL_REQUIRE(ct_supported(ctx, d), "okl_conv2_fprop: unsupported shape / layout");
  OKL_REQUIRE(d->act == OKL_ACT_NONE || d->act == OKL_ACT_RELU, "okl_conv2_fprop: only ReLU can be fused (act=%d)", d->act);
  CtShape s;
  ct_shape(d, s);
  const int vol3[4] = {s.D, s.KD, s.PD, s.OD};
  return conv_tc_run(ctx, x16, s.N, s.H, s.W, s.C, w2_16, s.O, s.KH, s.KW, s.PH, s.PW, s.OH, s.OW, (const float*)bias, d->act == OKL_ACT_RELU ? 1 : 0,
                     (float*)y, y16, nullptr, d->nd == 3 ? vol3 : nullptr);
}

extern "C" int okl_conv2_dgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* g16, const void* wf_16, void* dx, void* dx16,
                               const void* relu_mask16) {
  OKL_REQUIRE(ctx && d && g16 && wf_16 && (dx || dx16), "NULL argument");
  OKL_REQUIRE(ct_supported(ctx, d), "okl_conv2_dgrad: unsupported shape / layout");
  CtShape s;
  ct_shape(d, s);
  // dx = conv(g, flip(W)^T) with padding k - 1 - p: the same kernel with the roles of C and O exchanged
  const int vol3[4] = {s.OD, s.KD, s.KD - 1 - s.PD, s.D};     // the gradient's depth is the "input" depth of this convolution
  return conv_tc_run(ctx, g16, s.N, s.OH, s.OW, s.O, wf_16, s.C, s.KH, s.KW, s.KH - 1 - s.PH, s.KW - 1 - s.PW, s.H, s.W, nullptr, 0, (float*)dx, dx16,
                     relu_mask16, d->nd == 3 ? vol3 : nullptr);
}

// ================================================================================================ weight gradient
// dW[o][c][u][v] = sum_{n,h,w} g[n,h,w,o] * x[n, h+u-ph, w+v-pw, c]   (okl.py:739-741, 869-871: filter gradient of Conv1D / Conv2D)
//
// GEMM per CTA: D_u[(v, c) (M = 128), o (N = BN)] += x_slab_u^T . g over the CTA's share of 128-pixel tiles (split-K over CTAs, fixed-
// order second-stage reduce).  Both operands are MN-major straight from the channels-last bf16 tensors: one TMA box of
// (TH + KH - 1) input rows per (v, 64-channel chunk) - the KH vertical taps are the same slab at a row offset, each with its own
// TMEM accumulator - and one box of the gradient tile per 64 output channels, shared by every tap of the CTA.  An M tile is two
// 64-channel chunks; a chunk slot without data (C = 64 layers: 3 taps = 3 chunks) reads a constant all-ones region, which makes
// row 64 of that accumulator the bias gradient for free.  The first generation re-read the gradient once per tap and N tile
// (3.6 GB of L2 -> SM traffic on the 64 -> 64 layer at batch 1024, ncu) and was bound by exactly that.
namespace {

constexpr int WG_THREADS = 256;        // warp 0: TMA producer, 1: MMA issuer, 2: TMEM allocator, 4-7: epilogue
constexpr int WG_SLOT = 24576;         // one activation slab: (TH + KH - 1) rows of a 128-pixel tile, 64 channels
constexpr int WG_GCH = 16384;          // one gradient chunk: 128 pixels x 64 channels
constexpr int WG_ONES = 32768;         // constant region of bf16 1.0 (covers the largest shifted 128-pixel view)
constexpr int WG_PT = 128;             // pixels per k-block

struct WgParams {
  int tiles_mt, tiles_n;               // work tiles: M-tile groups x N tiles; splits CTAs each
  int splits, ktiles;                  // pixel tiles in total
  int tiles_w, tiles_h;
  int TW, TH, TN, tw_sh, pimg_sh;      // pixel tile TW x TH x TN = 128; pimg_sh = log2(TW * TH)
  int KH, KW, cpc, nchunks;            // cpc = C / 64; nchunks = KW * cpc real (v, channel-chunk) slots
  int MT;                              // M tiles (of two chunk slots) per CTA: 1 or 2
  int ph, pw;
  int slab_bytes, row_bytes, img_bytes;
  int stages, stage_bytes, nreal_max;
  int BN, has_db;
  float* ws;                           // [split][tile][MT * KH][128][BN]
};

template <int BN, int KHT>               // KHT: vertical taps known at compile time (3: the issue loop is fully unrolled), 0: run time
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sSt = smem;
  uint8_t* sOnes = sSt + p.stages * p.stage_bytes;          // after the stages: "ones - slab" is a positive leading-dimension offset
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + WG_ONES);
  uint64_t* full = bars;
  uint64_t* empty = bars + 4;
  uint64_t* done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < WG_ONES / 16; i += WG_THREADS) reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  okl_pdl_sync();

  // work decomposition: blockIdx.x = (tile, split)
  const int tile = blockIdx.x / p.splits, split = blockIdx.x - tile * p.splits;
  const int tmt = tile / p.tiles_n, tn = tile - tmt * p.tiles_n;
  const int q0 = tmt * p.MT * 2;                           // first chunk slot (flattened (v, channel chunk)) of this CTA
  int nreal = p.nchunks - q0;
  if (nreal > p.MT * 2) nreal = p.MT * 2;
  const int gbytes = (BN / 64) * WG_GCH;
  const uint32_t tx_bytes = (uint32_t)(nreal * p.slab_bytes + gbytes);

  if (warp == 0) {
    if (elect_one()) {
      int st = 0;
      uint32_t ph_ = 0;
      for (int kt = split; kt < p.ktiles; kt += p.splits) {
        const int wt = kt % p.tiles_w, r2 = kt / p.tiles_w;
        const int ht = r2 % p.tiles_h, ng = r2 / p.tiles_h;
        const int w0 = wt * p.TW, h0 = ht * p.TH, n0 = ng * p.TN;
        mbar_wait(&empty[st], ph_ ^ 1);
        mbar_expect_tx(&full[st], tx_bytes);
        uint8_t* base = sSt + st * p.stage_bytes;
        for (int j = 0; j < nreal; j++) {
          const int q = q0 + j, v = q / p.cpc, cc = q - v * p.cpc;
          tma_load_4d(&tmX, &full[st], base + j * WG_SLOT, cc * 64, w0 + v - p.pw, h0 - p.ph, n0);
        }
        uint8_t* gb = base + p.nreal_max * WG_SLOT;
#pragma unroll
        for (int j = 0; j < BN / 64; j++) tma_load_4d(&tmG, &full[st], gb + j * WG_GCH, tn * BN + j * 64, w0, h0, n0);
        if (++st == p.stages) { st = 0; ph_ ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // Single-thread issue loop, kept to one 32-bit add per MMA: everything of a descriptor that does not change with the 16-pixel group
      // is folded into one low word per (M tile, vertical tap) and stage, the per-group offsets are computed once.  (First version: a
      // triple loop that rebuilt 64-bit descriptors - 14 dependent uniform-datapath instructions, ~100 cycles, per 32-cycle MMA; issuing
      // every MMA twice cost 17 %.)
      constexpr uint32_t idesc = make_idesc(1, true, true, 128, BN);
      constexpr uint32_t HI = TC_DESC_HI_SW128;
      const int KH = KHT ? KHT : p.KH;
      const uint32_t ones16 = (smem_u32(sOnes) & 0x3FFFFu) >> 4;
      const uint32_t st0_16 = (smem_u32(sSt) & 0x3FFFFu) >> 4, stage16 = (uint32_t)p.stage_bytes >> 4, row16 = (uint32_t)p.row_bytes >> 4;
      const int nt = (nreal + p.has_db + 1) >> 1 < p.MT ? (nreal + p.has_db + 1) >> 1 : p.MT;     // M tiles with anything in them
      uint32_t off16[WG_PT / 16];
#pragma unroll
      for (int kk = 0; kk < WG_PT / 16; kk++) {
        const int pix = kk * 16;
        const int img = pix >> p.pimg_sh, pin = pix & ((1 << p.pimg_sh) - 1);
        off16[kk] = (uint32_t)(img * p.img_bytes + (pin >> p.tw_sh) * p.row_bytes + (pin & (p.TW - 1)) * 128) >> 4;
      }
      // M tile t = chunk slots 2t, 2t + 1: two real slabs (LBO = slot distance), a real slab + the constant ones region (LBO depends on
      // the stage), or ones only (absolute address)
      uint32_t rel[2];
      int kind[2];                                         // 0: two slabs, 1: slab + ones, 2: ones only
#pragma unroll
      for (int t = 0; t < 2; t++) {
        const bool r0 = 2 * t < nreal, r1 = 2 * t + 1 < nreal;
        kind[t] = r0 ? (r1 ? 0 : 1) : 2;
        rel[t] = kind[t] == 0 ? (((uint32_t)(2 * t * WG_SLOT) >> 4) | ((uint32_t)(WG_SLOT >> 4) << 16))
               : kind[t] == 1 ? ((uint32_t)(2 * t * WG_SLOT) >> 4)
                              : (ones16 | ((1024u >> 4) << 16));
      }
      const uint32_t g_rel = ((uint32_t)(p.nreal_max * WG_SLOT) >> 4) | ((uint32_t)(WG_GCH >> 4) << 16);
      const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
      uint32_t st = 0, ph_ = 0, accum = 0;
#pragma unroll 1
      for (int kt = split; kt < p.ktiles; kt += p.splits) {
        const uint32_t base16 = st0_16 + st * stage16;
        uint32_t lo[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
          lo[t] = kind[t] == 2 ? rel[t] : base16 + rel[t];
          if (kind[t] == 1) lo[t] |= (ones16 - lo[t]) << 16;                     // second chunk slot = the ones region
        }
        const uint32_t g_lo = base16 + g_rel;
        mbar_wait_a(full0 + st * 8, ph_);
        if (KHT) {
#pragma unroll
          for (int kk = 0; kk < WG_PT / 16; kk++) {
#pragma unroll
            for (int t = 0; t < 2; t++) {
              if (t < nt) {
#pragma unroll
                for (int u = 0; u < (KHT ? KHT : 1); u++)
                  umma_bf16_words<0>(tmem_base + (uint32_t)((t * KHT + u) * BN), lo[t] + off16[kk] + (uint32_t)u * row16, g_lo + (uint32_t)(kk * 128), HI,
                                     idesc, kk ? 1u : accum);
              }
            }
          }
        } else {
#pragma unroll 1
          for (int kk = 0; kk < WG_PT / 16; kk++) {
            const uint32_t gl = g_lo + (uint32_t)(kk * 128);
            uint32_t d = tmem_base;
            for (int t = 0; t < nt; t++) {
              uint32_t al = lo[t] + off16[kk];
              for (int u = 0; u < KH; u++, al += row16, d += BN) umma_bf16_words<0>(d, al, gl, HI, idesc, accum);
            }
            accum = 1;
          }
        }
        accum = 1;
        umma_commit_a(empty0 + st * 8);
        if (++st == (uint32_t)p.stages) { st = 0; ph_ ^= 1; }
      }
      umma_commit(done);
    }
  }
  // ===================================================================== epilogue (all eight warps): accumulators -> fp32 partials
  // As in the gradient-slab kernel below: scattered 16-byte stores straight from the TMEM registers ran at ~15 GB/s per SM (12.6 us per
  // 196 KB, probe), so every 128 x BN accumulator image is staged in the idle operand stages + ones region - segment (32-column chunk
  // c, row r) at (c * 128 + r) * 128 B, 16-byte pieces XOR-swizzled with the row - and leaves with one bulk copy; the second stage
  // reads the same swizzled image.
  __syncwarp();
  {
    const int q = warp & 3, par = warp >> 2;
    mbar_wait(done, 0);
    tc_fence_after();
    const bool any = split < p.ktiles;
    const int nacc = p.MT * p.KH, nch = BN >> 5;
    const uint32_t img_bytes = (uint32_t)(128 * BN * 4);
    const int nbuf = (int)(((uint32_t)(p.stages * p.stage_bytes) + (uint32_t)WG_ONES) / img_bytes);      // >= 1 (checked by the plan)
    float* dst0 = p.ws + (((size_t)split * (p.tiles_mt * p.tiles_n) + tile) * nacc) * (size_t)(128 * BN);
    const int r = q * 32 + lane;
    int na = 0;
    for (int a = 0; a < nacc; a++)
      if (2 * (a / p.KH) < nreal + p.has_db) na = a + 1;   // accumulators of the M tiles the MMA loop touched
    for (int a = 0; a < na; a++) {
      uint8_t* img = sSt + (size_t)(a % nbuf) * img_bytes;
      if (a >= nbuf) {                                     // the copy that last used this buffer must have read it
        if (threadIdx.x == 0) tma_store_wait_read<0>();
        __syncthreads();
      }
#pragma unroll 1
      for (int c = par; c < nch; c += 2) {
        uint32_t v[32];
        if (any) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c * 32), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = 0u;
        }
        uint4* seg = reinterpret_cast<uint4*>(img + ((size_t)(c * 128 + r) << 7));
#pragma unroll
        for (int j = 0; j < 8; j++) seg[j ^ (r & 7)] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      fence_proxy_async();
      __syncthreads();
      if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst0 + (size_t)a * 128 * BN), "r"(smem_u32(img)), "r"(img_bytes)
                     : "memory");
        tma_store_commit();
      }
    }
    if (threadIdx.x == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// dW[o][c][u][v] = sum_split ws[split][tile][t * KH + u][m][o'] ; db[o] from the ones slot.  One thread per (q, m64, u, o): consecutive
// threads walk o (the contiguous dimension of the partials).
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dW, float* __restrict__ db, int O, int C, int KH, int KW,
                                         int cpc, int nchunks, int MT, int BN, int tiles_mt, int tiles_n, int splits) {
  // four threads per output element (adjacent lanes): each sums every fourth split with four loads in flight, then a fixed-order
  // two-step shuffle - the loop is pure load latency, so the parallelism is what sets its time
  const int nq = nchunks + (db ? 1 : 0);
  const size_t total = (size_t)nq * 64 * KH * O;
  const int nacc = MT * KH;
  const size_t gtid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const int part = (int)(gtid & 3);
  for (size_t i = gtid >> 2; i < ((total + 7) & ~size_t(7)); i += ((size_t)gridDim.x * blockDim.x) >> 2) {
    const bool live = i < total;
    const size_t ii = live ? i : 0;
    const int o = (int)(ii % O);
    size_t r = ii / O;
    const int u = (int)(r % KH); r /= KH;
    const int m64 = (int)(r % 64);
    const int q = (int)(r / 64);
    const bool is_db = q == nchunks;
    const int tmt = q / (2 * MT), slot = q - tmt * 2 * MT, t = slot >> 1, j = slot & 1;
    const int tn = o / BN, on = o - tn * BN;
    const int tile = tmt * tiles_n + tn;
    const size_t stride = (size_t)(tiles_mt * tiles_n) * nacc * 128 * BN;
    const int row = j * 64 + m64;                           // swizzled segment image written by the kernel's epilogue
    const float* src = ws + ((size_t)tile * nacc + (t * KH + u)) * (size_t)(128 * BN) + ((size_t)((on >> 5) * 128 + row) << 5) +
                       (((((on & 31) >> 2) ^ (row & 7)) << 2) | (on & 3));
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = part;
    for (; k + 12 < splits; k += 16) {
      const float t0 = __ldcs(src + (size_t)k * stride), t1 = __ldcs(src + (size_t)(k + 4) * stride);
      const float t2 = __ldcs(src + (size_t)(k + 8) * stride), t3 = __ldcs(src + (size_t)(k + 12) * stride);
      a0 += t0; a1 += t1; a2 += t2; a3 += t3;
    }
    for (; k < splits; k += 4) a0 += __ldcs(src + (size_t)k * stride);
    float s = (a0 + a1) + (a2 + a3);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (!live || part != 0) continue;
    if (is_db) {
      if (m64 == 0 && u == 0) db[o] = s;
    } else {
      const int v = q / cpc, c = (q - v * cpc) * 64 + m64;
      dW[(((size_t)o * C + c) * KH + u) * KW + v] = s;
    }
  }
}

struct WgPlan { WgParams p; int smem; bool db_fused; };

bool wg_plan(okl_ctx* ctx, const okl_conv_desc* d, bool want_db, WgPlan& pl) {
  if (!ct_supported(ctx, d) || ct_env("OKL_CONV_TC2_WGRAD", 1) == 0) return false;
  CtShape s;
  ct_shape(d, s);
  // measured (batch 1024): 64-channel inputs 197 vs 321 us and 77 vs 91 us against the first-generation kernel; for C >= 128 the
  // first generation's N = 128 / 256 tiles still win (98 / 59 / 81 us vs 137 / 96 / 133 us) - those shapes stay there for now
  if (s.C > 64 && ct_env("OKL_WG_ALL", 0) == 0) return false;
  if (s.C % 64 || s.O % 64 || d->nd == 3) return false;    // whole 64-channel chunks, 1-D / 2-D only (the gradient-slab kernel takes the rest)
  WgParams& p = pl.p;
  memset(&p, 0, sizeof(p));
  p.TW = ct_pow2_ceil(s.OW); if (p.TW > 32) p.TW = 32;
  if (p.TW < 8) return false;
  p.TH = ct_pow2_ceil(s.OH); if (p.TH > WG_PT / p.TW) p.TH = WG_PT / p.TW;
  p.TN = WG_PT / (p.TW * p.TH);
  if (p.TW * p.TH < 16) return false;
  p.tw_sh = ct_ilog2(p.TW); p.pimg_sh = ct_ilog2(p.TW * p.TH);
  p.KH = s.KH; p.KW = s.KW; p.cpc = s.C / 64; p.nchunks = s.KW * p.cpc; p.ph = s.PH; p.pw = s.PW;
  p.row_bytes = p.TW * 128;
  p.img_bytes = (p.TH + s.KH - 1) * p.row_bytes;
  p.slab_bytes = p.TN * p.img_bytes;
  if (p.slab_bytes > WG_SLOT) return false;
  p.BN = s.O % 128 == 0 ? 128 : 64;
  if (s.KH * p.BN > 512) p.BN = 64;
  if (s.KH * p.BN > 512) return false;
  p.MT = (2 * s.KH * p.BN <= 512 && p.nchunks > 2 && ct_env("OKL_WG_MT", 2) >= 2) ? 2 : 1;
  const int slots = p.MT * 2;
  p.tiles_mt = (p.nchunks + (want_db ? 1 : 0) + slots - 1) / slots;
  // the ones slot gives db only when a chunk slot is free anyway; otherwise db is left to the caller
  pl.db_fused = want_db && p.tiles_mt == (p.nchunks + slots - 1) / slots;
  if (!pl.db_fused) p.tiles_mt = (p.nchunks + slots - 1) / slots;
  p.has_db = pl.db_fused ? 1 : 0;
  p.tiles_n = s.O / p.BN;
  p.nreal_max = p.nchunks < slots ? p.nchunks : slots;
  p.stage_bytes = p.nreal_max * WG_SLOT + (p.BN / 64) * WG_GCH;
  p.stages = (225 * 1024 - WG_ONES - 1024 - 256) / p.stage_bytes;
  if (p.stages > 4) p.stages = 4;
  if (p.stages < 2) return false;
  pl.smem = WG_ONES + p.stages * p.stage_bytes + 256 + 1024;
  p.tiles_w = (s.OW + p.TW - 1) / p.TW; p.tiles_h = (s.OH + p.TH - 1) / p.TH;
  p.ktiles = p.tiles_w * p.tiles_h * ((s.N + p.TN - 1) / p.TN);
  const int tiles = p.tiles_mt * p.tiles_n;
  p.splits = ctx->sm_count / tiles;
  if (p.splits > p.ktiles) p.splits = p.ktiles;
  if (p.splits < 1) p.splits = 1;
  return true;
}

template <int BN, int KHT>
int wg_launch(okl_ctx* ctx, const CUtensorMap& tx, const CUtensorMap& tg, const WgPlan& pl) {
  auto kern = conv_wgrad_tc_kernel<BN, KHT>;
  static int attr_smem = 0;
  if (attr_smem < pl.smem) {
    OKL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_smem = 227 * 1024;
  }
  okl_launch(ctx, kern, dim3(pl.p.tiles_mt * pl.p.tiles_n * pl.p.splits), dim3(WG_THREADS), pl.smem, tx, tg, pl.p);
  OKL_LAUNCHED(ctx, "conv_wgrad_tc_kernel");
  return OKL_OK;
}


// ------------------------------------------------------------------------------------------------ weight gradient, taps stacked along N
// The kernel above issues one 128 x 64 (or 128) x 16 MMA per (M tile, vertical tap): at N = 64 an MMA fetches (128 + 64) x 32 B of
// operands for 32 cycles of tensor work - 192 B/cycle against the SM's 128 B/cycle of shared memory - and a quarter of the M rows of
// the 64-channel layers is the ones slot: 43 % of the tensor peak measured, 50 % possible.  Here the vertical taps move to the OTHER
// operand.  With h' = h + u - ph (an INPUT row):
//     dW[o][c][u][v] = sum_{n,h',w} x[n, h', w + v - pw, c] * g[n, h' - u + ph, w, o]
// so a k-block is a tile of 128 (or 64) input pixels: A = one plain x box per (v, 64-channel chunk) (M = 128 = two chunk slots), and
// B = ONE gradient slab of TH + KH - 1 rows whose KH row-shifted views are the KH 64-column groups of a single N = 64 KH operand: the
// leading-dimension offset of the MN-major descriptor (the distance between 64-element groups along N) is simply one slab row.  One
// 128 x 192 x 16 MMA replaces three 128 x 64 x 16 ones: (128 + 192) x 32 B per 96 cycles = 107 B/cycle.  TMA's zero fill supplies the
// gradient rows outside the image.  Column group j of an accumulator is tap u = KH - 1 - j.  A CTA owns MT M tiles x NO 64-channel
// output chunks (MT x NO <= 2 accumulators of 256 TMEM columns); split-K over the SMs, fixed-order second stage as before.  A free
// chunk slot reads the all-ones region: row 64 of that accumulator, column group u = ph, is the bias gradient.
constexpr int GS_THREADS = 256;        // warp 0: TMA producer, 1: MMA issuer, 2: TMEM allocator, 4-7: epilogue
constexpr int GS_MAXST = 6;

struct GsParams {
  int tiles_mt, tiles_n, splits, ktiles;
  int tiles_w, tiles_h;
  int PT, TW, TH, TN, tw_sh, pimg_sh;  // k-block = TW x TH pixels of TN images = PT (64 or 128)
  int KH, KW, cpc, nchunks;            // cpc = C / 64; nchunks = KW * cpc real (v, channel chunk) slots
  int MT, NO;                          // per CTA: M tiles (two chunk slots each), 64-channel output chunks
  int ph, pw;
  int xbox_bytes, gslab_bytes, row_bytes, gimg_bytes;
  int stages, stage_bytes, nreal_max;
  int has_db, ones_bytes;
  int KD, pd, D;                       // Conv3D: depth taps (one per CTA group: tile = (depth tap, M tile group, N tile)), depth padding, input
                                       // depth; D = 0 for 1-D / 2-D layers.  An image of the k-tile decode is one (sample, input depth) slice.
  int epi_bufs;                        // staging images that fit the operand stages (1 or 2)
  int diag;                            // probe only (OKL_WG_DIAG): 1 = every MMA issued twice, 2 = operands loaded once per stage slot
  float* ws;                           // [split][tile][MT * NO][128][64 KH]
};

__global__ void __launch_bounds__(GS_THREADS, 1)
conv_wgrad_gs_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const GsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sSt = smem;
  uint8_t* sOnes = sSt + p.stages * p.stage_bytes;          // after the stages: "ones - box" is a positive leading-dimension offset
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + p.ones_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + GS_MAXST;
  uint64_t* done = bars + 2 * GS_MAXST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GS_MAXST + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < p.ones_bytes / 16; i += GS_THREADS)
    reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  okl_pdl_sync();

  const int tile = blockIdx.x / p.splits, split = blockIdx.x - tile * p.splits;
  const int tt = tile / (p.tiles_mt * p.tiles_n), tile2 = tile - tt * (p.tiles_mt * p.tiles_n);        // depth tap of this CTA (0 for 2-D)
  const int tmt = tile2 / p.tiles_n, tn = tile2 - tmt * p.tiles_n;
  const int q0 = tmt * p.MT * 2;                           // first chunk slot (flattened (v, channel chunk)) of this CTA
  int nreal = p.nchunks - q0;
  if (nreal > p.MT * 2) nreal = p.MT * 2;
  const int nslots = nreal + ((p.has_db && q0 + nreal == p.nchunks && nreal < p.MT * 2) ? 1 : 0);      // + the ones slot
  const int nt = (nslots + 1) >> 1;                        // M tiles with anything in them
  const int NC = p.KH * 64;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx_bytes = (uint32_t)(nreal * p.xbox_bytes + p.NO * p.gslab_bytes);
      int st = 0;
      uint32_t ph_ = 0;
      for (int kt = split; kt < ((p.diag & 4) ? 0 : p.ktiles); kt += p.splits) {
        const int wt = kt % p.tiles_w, r2 = kt / p.tiles_w;
        const int ht = r2 % p.tiles_h, ng = r2 / p.tiles_h;
        const int w0 = wt * p.TW, h0 = ht * p.TH, n0 = ng * p.TN;
        mbar_wait(&empty[st], ph_ ^ 1);
        if ((p.diag & 2) && kt >= split + p.stages * p.splits) {
          mbar_arrive(&full[st]);
          if (++st == p.stages) { st = 0; ph_ ^= 1; }
          continue;
        }
        mbar_expect_tx(&full[st], tx_bytes);
        uint8_t* base = sSt + st * p.stage_bytes;
        // 3-D: image group -> (sample, first input depth); the gradient slab of depth tap tt sits tt - pd slices before it (zero fill outside)
        const int smp = p.D > 0 ? n0 / p.D : 0, d0 = p.D > 0 ? n0 - smp * p.D : 0;
        for (int j = 0; j < nreal; j++) {
          const int q = q0 + j, v = q / p.cpc, cc = q - v * p.cpc;
          if (p.D > 0) tma_load_5d(&tmX, &full[st], base + j * p.xbox_bytes, cc * 64, w0 + v - p.pw, h0, d0, smp);
          else tma_load_4d(&tmX, &full[st], base + j * p.xbox_bytes, cc * 64, w0 + v - p.pw, h0, n0);
        }
        uint8_t* gb = base + p.nreal_max * p.xbox_bytes;
        for (int j = 0; j < p.NO; j++) {
          if (p.D > 0) tma_load_5d(&tmG, &full[st], gb + j * p.gslab_bytes, (tn * p.NO + j) * 64, w0, h0 + p.ph - (p.KH - 1), d0 - tt + p.pd, smp);
          else tma_load_4d(&tmG, &full[st], gb + j * p.gslab_bytes, (tn * p.NO + j) * 64, w0, h0 + p.ph - (p.KH - 1), n0);
        }
        if (++st == p.stages) { st = 0; ph_ ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(1, true, true, 128, (p.diag & 32) ? 64 : NC);
      constexpr uint32_t HI = TC_DESC_HI_SW128;
      const uint32_t ones16 = (smem_u32(sOnes) & 0x3FFFFu) >> 4;
      const uint32_t st0_16 = (smem_u32(sSt) & 0x3FFFFu) >> 4, stage16 = (uint32_t)p.stage_bytes >> 4, xbox16 = (uint32_t)p.xbox_bytes >> 4;
      const uint32_t g0_16 = (uint32_t)(p.nreal_max * p.xbox_bytes) >> 4, gslab16 = (uint32_t)p.gslab_bytes >> 4;
      const uint32_t row_lbo = ((uint32_t)p.row_bytes >> 4) << 16;
      const int nk = p.PT >> 4, NO = p.NO;
      uint32_t goff16[8];
#pragma unroll
      for (int kk = 0; kk < 8; kk++) {
        const int pix = kk * 16;
        const int img = pix >> p.pimg_sh, pin = pix & ((1 << p.pimg_sh) - 1);
        goff16[kk] = (uint32_t)(img * p.gimg_bytes + (pin >> p.tw_sh) * p.row_bytes + (pin & (p.TW - 1)) * 128) >> 4;
      }
      // M tile t = chunk slots 2t, 2t + 1: two x boxes (LBO = box size), or a box + the ones region (LBO depends on the stage), or a box
      // alone (its second 64 rows repeat the first and are never read back)
      int kind[2];
#pragma unroll
      for (int t = 0; t < 2; t++) kind[t] = 2 * t + 1 < nreal ? 0 : (2 * t + 1 < nslots ? 1 : 2);
      const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
      uint32_t st = 0, ph_ = 0, accum = 0;
#pragma unroll 1
      for (int rep = 0; rep < ((p.diag & 256) ? 2 : 1); rep++)
#pragma unroll 1
      for (int kt = split; kt < p.ktiles; kt += p.splits) {
        const uint32_t base16 = st0_16 + st * stage16;
        uint32_t alo[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
          const uint32_t a0 = base16 + (uint32_t)(2 * t) * xbox16;
          alo[t] = a0 | ((kind[t] == 0 ? xbox16 : kind[t] == 1 ? ones16 - a0 : 0u) << 16);
          if (p.diag & 64) alo[t] = a0;
        }
        const uint32_t glo = (base16 + g0_16) | row_lbo;
        if (!(p.diag & 4)) mbar_wait_a(full0 + st * 8, ph_);
        tc_fence_after();
        const uint32_t dmul = (p.diag & 8) ? 0u : 256u;
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          if (kk < nk) {
            const uint32_t ak = (uint32_t)kk * (2048u >> 4), gk = goff16[kk];
#pragma unroll
            for (int t = 0; t < 2; t++)
#pragma unroll
              for (int j = 0; j < 2; j++)
                if (t < nt && j < NO) {
                  umma_bf16_words<0>(tmem_base + (uint32_t)(t * NO + j) * dmul, alo[t] + ak, glo + (uint32_t)j * gslab16 + gk, HI, idesc, accum);
                  if (p.diag & 1)
                    umma_bf16_words<0>(tmem_base + (uint32_t)(t * NO + j) * dmul, alo[t] + ak, glo + (uint32_t)j * gslab16 + gk, HI, idesc, 1u);
                }
            accum = 1;
          }
        }
        umma_commit_a(empty0 + st * 8);
        if (++st == (uint32_t)p.stages) { st = 0; ph_ ^= 1; }
      }
      umma_commit(done);
    }
  }
  // ===================================================================== epilogue (all eight warps): accumulators -> fp32 partials
  // Partials of a unit are one 128 x 64 KH fp32 image, contiguous in the workspace.  Scattered 16-byte stores straight from the TMEM
  // registers cost 12.6 us per CTA (probe); instead the image is staged in the (now idle) operand stages - segment (32-column chunk c,
  // row r) at ((c * 128 + r) * 128 B, its 16-byte pieces XOR-swizzled with the row so that a warp's stores hit all banks - and leaves
  // with ONE bulk copy per unit.  The second-stage kernel reads the same swizzled image.
  __syncwarp();
  {
    const int q = warp & 3, par = warp >> 2;               // TMEM lane quarter; chunks of parity par
    mbar_wait(done, 0);
    tc_fence_after();
    const int nunits = p.MT * p.NO, nch = NC >> 5;
    const uint32_t unit_bytes = (uint32_t)(128 * NC * 4);
    const int ntiles = (p.D > 0 ? p.KD : 1) * p.tiles_mt * p.tiles_n;
    float* dst0 = p.ws + (((size_t)split * ntiles + tile) * nunits) * (size_t)(128 * NC);
    const int r = q * 32 + lane;
    const int nu = (p.diag & 128) ? 0 : nt * p.NO;
    for (int a = 0; a < nu; a++) {
      uint8_t* img = sSt + (size_t)(a & (p.epi_bufs - 1)) * unit_bytes;
      if (a >= p.epi_bufs) {                               // single staging buffer: the previous unit's copy must have read it
        if (threadIdx.x == 0) tma_store_wait_read<0>();
        __syncthreads();
      }
#pragma unroll 1
      for (int c = par; c < nch; c += 2) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256 + c * 32), v);
        tmem_ld_wait();
        uint4* seg = reinterpret_cast<uint4*>(img + ((size_t)(c * 128 + r) << 7));
#pragma unroll
        for (int j = 0; j < 8; j++) seg[j ^ (r & 7)] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      fence_proxy_async();
      __syncthreads();
      if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst0 + (size_t)a * 128 * NC), "r"(smem_u32(img)), "r"(unit_bytes)
                     : "memory");
        tma_store_commit();
      }
    }
    if (threadIdx.x == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Second stage: dW[o][c][u][v] = sum over splits, in fixed order, of the partial images; db[o] from the ones slot (column group u = ph).
// A warp owns one 128-byte segment (tile, unit, row, 32-column chunk) and every G-th split of it (coalesced loads, all in flight); the G
// warps of a segment sit in one block, the first adds their sums in order and scatters into the reference's (O, C, KH, KW) layout.
__global__ void __launch_bounds__(256) conv_wgrad_gs_reduce(const float* __restrict__ ws, float* __restrict__ dW, float* __restrict__ db, int O, int C,
                                                            int KH, int KW, int cpc, int nchunks, int MT, int NO, int tiles_mt, int tiles_n,
                                                            int splits, int ph, int G, int nseg, int KD, int pd) {
  __shared__ float part[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NC = KH * 64, nch = NC >> 5, nunits = MT * NO;
  const int g = warp & (G - 1);
  int b = blockIdx.x * (8 / G) + warp / G;
  const bool in_range = b < nseg;
  if (!in_range) b = 0;
  const int c = b % nch; b /= nch;
  const int row = b & 127; b >>= 7;
  const int unit = b % nunits, tile = b / nunits;
  const int tt = tile / (tiles_mt * tiles_n), tile2 = tile - tt * (tiles_mt * tiles_n);      // depth tap (0 for 2-D)
  const int tmt = tile2 / tiles_n, tn = tile2 - tmt * tiles_n;
  const int t = unit / NO, jo = unit - t * NO;
  const int q = tmt * 2 * MT + 2 * t + (row >> 6), m64 = row & 63;      // chunk slot of this row
  const bool is_db = db && q == nchunks && m64 == 0 && tt == pd;
  const bool live = in_range && (q < nchunks || is_db);
  float sum = 0.f;
  if (live) {
    const size_t stride = (size_t)(KD * tiles_mt * tiles_n) * nunits * 128 * NC;
    const float* src = ws + ((size_t)tile * nunits + unit) * 128 * NC + ((size_t)(c * 128 + row) << 5) + ((((lane >> 2) ^ (row & 7)) << 2) | (lane & 3));
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = g;
    for (; k + 7 * G < splits; k += 8 * G) {
      float tv[8];
#pragma unroll
      for (int i = 0; i < 8; i++) tv[i] = __ldcs(src + (size_t)(k + i * G) * stride);
      a0 += tv[0]; a1 += tv[1]; a2 += tv[2]; a3 += tv[3];
      a0 += tv[4]; a1 += tv[5]; a2 += tv[6]; a3 += tv[7];
    }
    for (; k + 3 * G < splits; k += 4 * G) {
      const float t0 = __ldcs(src + (size_t)k * stride), t1 = __ldcs(src + (size_t)(k + G) * stride);
      const float t2 = __ldcs(src + (size_t)(k + 2 * G) * stride), t3 = __ldcs(src + (size_t)(k + 3 * G) * stride);
      a0 += t0; a1 += t1; a2 += t2; a3 += t3;
    }
    for (; k < splits; k += G) a0 += __ldcs(src + (size_t)k * stride);
    sum = (a0 + a1) + (a2 + a3);
  }
  part[warp][lane] = sum;
  __syncthreads();
  if (!live || g != 0) return;
  float s = 0.f;
  for (int w = 0; w < G; w++) s += part[warp + w][lane];
  const int col = c * 32 + lane, u = KH - 1 - (col >> 6), o = (tn * NO + jo) * 64 + (col & 63);
  if (o >= O) return;                                      // partial last 64-channel chunk
  if (q == nchunks) {
    if (u == ph) db[o] = s;
  } else {
    const int v = q / cpc, ch = (q - v * cpc) * 64 + m64;
    if (ch < C) dW[((((size_t)o * C + ch) * KD + tt) * KH + u) * KW + v] = s;
  }
}

struct GsPlan { GsParams p; int smem; bool db_fused; };

bool gs_plan(okl_ctx* ctx, const okl_conv_desc* d, bool want_db, GsPlan& pl) {
  if (!ct_supported(ctx, d) || ct_env("OKL_CONV_TC2_WGRAD", 1) == 0 || ct_env("OKL_WG_GS", 1) == 0) return false;
  CtShape s;
  ct_shape(d, s);
  if (s.KH < 2 || s.KH > 4) return false;                  // N = 64 KH must be a real gain and fit one MMA
  // measured (batch 1024, whole call incl. second stage, us): 64->64@32 98 vs 105 (with db) / 93 vs 104 (db from the pooling backward);
  // 64->128@16 61 vs 60; C >= 128: 84 / 58 / 87 (+ channel sum) vs the first generation's 128 x 256 tiles 94 / 62 / 95 incl. db - its
  // N = 256 MMAs fetch fewer operand bytes per flop.  So: C = O = 64 only, unless OKL_WG_ALL asks for every shape (parity tests, probes).
  const bool partial_chunks = s.C % 64 || s.O % 64;        // only this kernel handles those (zero-filled channel boxes, guarded second stage)
  const bool vol = d->nd == 3;                             // ... and Conv3D (okl.py:1018-1031): one depth tap per CTA group
  if ((s.C != 64 || s.O != 64) && !partial_chunks && !vol && ct_env("OKL_WG_ALL", 0) == 0) return false;
  GsParams& p = pl.p;
  memset(&p, 0, sizeof(p));
  p.KD = vol ? s.KD : 1; p.pd = s.PD; p.D = vol ? s.D : 0;
  p.KH = s.KH; p.KW = s.KW; p.cpc = (s.C + 63) / 64; p.nchunks = s.KW * p.cpc; p.ph = s.PH; p.pw = s.PW;
  const int ochunks = (s.O + 63) / 64;
  // bias gradient for free when a chunk slot is empty anyway and column group u = ph sees every gradient row (H >= OH)
  pl.db_fused = want_db && (p.nchunks & 1) && 2 * s.PH <= s.KH - 1 && 2 * s.PD <= s.KD - 1;
  p.has_db = pl.db_fused ? 1 : 0;
  const int S = p.nchunks + p.has_db;
  if (S > 2 && S <= 4) { p.MT = 2; p.NO = 1; }
  else if (ochunks % 2 == 0) { p.MT = 1; p.NO = 2; }
  else { p.MT = 1; p.NO = 1; }
  p.tiles_mt = (S + 2 * p.MT - 1) / (2 * p.MT);
  p.tiles_n = ochunks / p.NO;
  p.nreal_max = p.nchunks < 2 * p.MT ? p.nchunks : 2 * p.MT;
  p.ones_bytes = p.has_db ? 16384 : 0;
  int pt_env = ct_env("OKL_WG_PT", 0);
  bool ok = false;
  for (int pt = 128; pt >= 64 && !ok; pt >>= 1) {
    if (pt_env && pt != pt_env) continue;
    p.PT = pt;
    p.TW = ct_pow2_ceil(s.OW); if (p.TW > 32) p.TW = 32;
    if (p.TW < 8) return false;
    p.TH = ct_pow2_ceil(s.H); if (p.TH > pt / p.TW) p.TH = pt / p.TW;
    p.TN = pt / (p.TW * p.TH);
    if (p.TW * p.TH < 16) continue;
    if (vol && s.D % p.TN) continue;                       // the images of a k-block are depth slices of one sample
    p.tw_sh = ct_ilog2(p.TW); p.pimg_sh = ct_ilog2(p.TW * p.TH);
    p.row_bytes = p.TW * 128;
    p.gimg_bytes = (p.TH + s.KH - 1) * p.row_bytes;
    p.gslab_bytes = p.TN * p.gimg_bytes;
    p.xbox_bytes = pt * 128;
    if (p.TH + s.KH - 1 > 256) continue;
    p.stage_bytes = p.nreal_max * p.xbox_bytes + p.NO * p.gslab_bytes;
    p.stages = (226 * 1024 - p.ones_bytes - 1024 - 256) / p.stage_bytes;
    if (p.stages > GS_MAXST) p.stages = GS_MAXST;
    const int want = ct_env("OKL_WG_MINST", 2);
    if (p.stages >= want || (pt == 64 && p.stages >= 2) || (pt_env && p.stages >= 2)) ok = true;
  }
  if (!ok) return false;
  pl.smem = p.ones_bytes + p.stages * p.stage_bytes + 256 + 1024;
  p.diag = ct_env("OKL_WG_DIAG", 0);
  p.epi_bufs = (p.MT * p.NO > 1 && p.stages * p.stage_bytes >= 2 * 128 * s.KH * 64 * 4) ? 2 : 1;
  if (p.stages * p.stage_bytes < 128 * s.KH * 64 * 4) return false;
  p.tiles_w = (s.OW + p.TW - 1) / p.TW; p.tiles_h = (s.H + p.TH - 1) / p.TH;
  p.ktiles = p.tiles_w * p.tiles_h * ((s.N * s.D + p.TN - 1) / p.TN);
  const int tiles = p.KD * p.tiles_mt * p.tiles_n;
  p.splits = ctx->sm_count / tiles;
  if (p.splits > p.ktiles) p.splits = p.ktiles;
  if (p.splits < 1) p.splits = 1;
  return true;
}

// shapes with partial 64-channel chunks: this kernel must take them, and the bias gradient must come either from its ones slot or from
// okl_channel_sum_bf16 (channel counts of 8 x a power of two)
bool gs_feasible(okl_ctx* ctx, const okl_conv_desc* d) {
  if (d->nd > 3 || d->x_layout != OKL_NXC || d->y_layout != OKL_NXC || d->C % 8 || d->O % 8) return false;
  CtShape s;
  ct_shape(d, s);
  if (s.KH < 2 || s.KH > 4 || s.OH < 1 || s.OW < 1) return false;
  if (ct_env("OKL_CONV_TC2_WGRAD", 1) == 0 || ct_env("OKL_WG_GS", 1) == 0) return false;
  const int nchunks = s.KW * ((s.C + 63) / 64);
  const bool db_fused = (nchunks & 1) && 2 * s.PH <= s.KH - 1 && 2 * s.PD <= s.KD - 1;
  if (d->nd == 3) {                                        // k-blocks of 64 pixels at least: their images must be slices of one sample
    const int tw3 = ct_pow2_ceil(s.OW) > 32 ? 32 : ct_pow2_ceil(s.OW);
    int th3 = ct_pow2_ceil(s.H); if (th3 > 64 / tw3) th3 = 64 / tw3 > 0 ? 64 / tw3 : 1;
    const int tn3 = 64 / (tw3 * th3) > 0 ? 64 / (tw3 * th3) : 1;
    if (s.D % tn3) return false;
  }
  const int o8 = s.O / 8;
  if (!db_fused && !((o8 & (o8 - 1)) == 0 && o8 <= 256)) return false;
  const int tw = ct_pow2_ceil(s.OW) > 32 ? 32 : ct_pow2_ceil(s.OW);
  if (tw < 8 || tw * ct_pow2_ceil(s.H) < 16) return false;
  return true;
}

int gs_run(okl_ctx* ctx, const okl_conv_desc* d, GsPlan& pl, const void* x16, const void* g16, void* dW, void* db) {
  CtShape s;
  ct_shape(d, s);
  GsParams& p = pl.p;
  const size_t ws_bytes = (size_t)p.splits * p.KD * p.tiles_mt * p.tiles_n * p.MT * p.NO * 128 * (p.KH * 64) * sizeof(float);
  int st = okl_ensure_workspace(ctx, ws_bytes);
  if (st != OKL_OK) return st;
  p.ws = (float*)ctx->workspace;
  CUtensorMap tx, tg;
  st = p.D > 0 ? ct_map_ndhwc(&tx, x16, s.C, s.W, s.H, s.D, s.N, 64, p.TW, p.TH, p.TN) : ct_map_nhwc(&tx, x16, 2, s.C, s.W, s.H, s.N, 64, p.TW, p.TH, p.TN);
  if (st != OKL_OK) return st;
  st = p.D > 0 ? ct_map_ndhwc(&tg, g16, s.O, s.OW, s.OH, s.OD, s.N, 64, p.TW, p.TH + s.KH - 1, p.TN)
               : ct_map_nhwc(&tg, g16, 2, s.O, s.OW, s.OH, s.N, 64, p.TW, p.TH + s.KH - 1, p.TN);
  if (st != OKL_OK) return st;
  static int attr_smem = 0;
  if (attr_smem < pl.smem) {
    OKL_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_gs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_smem = 227 * 1024;
  }
  okl_launch(ctx, conv_wgrad_gs_kernel, dim3(p.KD * p.tiles_mt * p.tiles_n * p.splits), dim3(GS_THREADS), pl.smem, tx, tg, p);
  OKL_LAUNCHED(ctx, "conv_wgrad_gs_kernel");
  const int nseg = p.KD * p.tiles_mt * p.tiles_n * p.MT * p.NO * 128 * (s.KH * 2);
  const int G = p.splits >= 32 ? 8 : (p.splits >= 16 ? 4 : (p.splits >= 8 ? 2 : 1));
  conv_wgrad_gs_reduce<<<(nseg + 8 / G - 1) / (8 / G), 256, 0, ctx->stream>>>(p.ws, (float*)dW, pl.db_fused ? (float*)db : nullptr, s.O, s.C, s.KH, s.KW,
                                                                                 p.cpc, p.nchunks, p.MT, p.NO, p.tiles_mt, p.tiles_n, p.splits, p.ph, G, nseg, p.KD, p.pd);
  OKL_LAUNCHED(ctx, "conv_wgrad_gs_reduce");
  return OKL_OK;
}

}  // namespace

extern "C" int okl_channel_sum_bf16(okl_ctx* ctx, const void* g16, void* out, long long rows, int C);

extern "C" int okl_conv2_wgrad_supported(okl_ctx* ctx, const okl_conv_desc* d) {
  if (!ctx || !d) return 0;
  GsPlan gp;
  if (gs_plan(ctx, d, false, gp)) return 1;
  WgPlan pl;
  return wg_plan(ctx, d, false, pl) ? 1 : 0;
}

// dW (fp32, reference layout (O, C, *k)); db (fp32 [O], optional): produced by the same kernel when a chunk slot is free, else by
// okl_channel_sum_bf16 on g16
extern "C" int okl_conv2_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* x16, const void* g16, void* dW, void* db) {
  OKL_REQUIRE(ctx && d && x16 && g16 && dW, "NULL argument");
  CtShape s;
  ct_shape(d, s);
  GsPlan gp;
  if (gs_plan(ctx, d, db != nullptr, gp)) {
    int st = gs_run(ctx, d, gp, x16, g16, dW, db);
    if (st != OKL_OK) return st;
    if (db && !gp.db_fused) return okl_channel_sum_bf16(ctx, g16, db, (long long)s.N * s.OD * s.OH * s.OW, s.O);
    return OKL_OK;
  }
  WgPlan pl;
  OKL_REQUIRE(wg_plan(ctx, d, db != nullptr, pl), "okl_conv2_wgrad: unsupported shape / layout");
  WgParams& p = pl.p;
  const size_t ws_bytes = (size_t)p.splits * p.tiles_mt * p.tiles_n * p.MT * p.KH * 128 * p.BN * sizeof(float);
  int st = okl_ensure_workspace(ctx, ws_bytes);
  if (st != OKL_OK) return st;
  p.ws = (float*)ctx->workspace;
  CUtensorMap tx, tg;
  st = ct_map_nhwc(&tx, x16, 2, s.C, s.W, s.H, s.N, 64, p.TW, p.TH + s.KH - 1, p.TN);
  if (st != OKL_OK) return st;
  st = ct_map_nhwc(&tg, g16, 2, s.O, s.OW, s.OH, s.N, 64, p.TW, p.TH, p.TN);
  if (st != OKL_OK) return st;
  if (s.KH == 3) st = p.BN == 128 ? wg_launch<128, 3>(ctx, tx, tg, pl) : wg_launch<64, 3>(ctx, tx, tg, pl);
  else st = p.BN == 128 ? wg_launch<128, 0>(ctx, tx, tg, pl) : wg_launch<64, 0>(ctx, tx, tg, pl);
  if (st != OKL_OK) return st;
  const size_t total = (size_t)(p.nchunks + (pl.db_fused ? 1 : 0)) * 64 * s.KH * s.O;
  conv_wgrad_reduce_kernel<<<okl_grid_1d(ctx, total * 4, 256, 16), 256, 0, ctx->stream>>>(p.ws, (float*)dW, pl.db_fused ? (float*)db : nullptr, s.O, s.C, s.KH,
                                                                                 s.KW, p.cpc, p.nchunks, p.MT, p.BN, p.tiles_mt, p.tiles_n, p.splits);
  OKL_LAUNCHED(ctx, "conv_wgrad_reduce");
  if (db && !pl.db_fused) return okl_channel_sum_bf16(ctx, g16, db, (long long)s.N * s.OH * s.OW, s.O);
  return OKL_OK;
}

// ================================================================================================ first layer (tiny C): K = C * taps < 128
// Conv2DLayer(3, 64, 3, 1, 1) on the input images (okl.py:790-813 / 869-871) and Conv3DLayer(3, 64, 3, 1, 1) on video clips
// (okl.py:948-985 / 1018-1031): K = 27 / 81 is far too small for a TMA-fed implicit GEMM (a k-block is 64 channels of ONE tap) and
// the kernels these layers used to run on took 16 % of the CIFAR step (FFMA: 202 us forward at 1.3 TB/s of fp32 output, 400 us weight
// gradient) and 0.86 of the 1.49 ms video step (sc_fprop / sc_wgrad: the im2col operand gathered tap by tap from L1 / L2).  Here the
// im2col operand of a 128-pixel tile - KP = 16 * KS bf16 per pixel: the taps, a constant 1 (bias / bias gradient ride in the same
// product) and zero padding - is BUILT in shared memory by the CTA's threads from a staged fp32 input patch (every input value is read
// from global memory once per tile), directly in the 128-byte-swizzled UMMA layout, and consumed by tcgen05.mma:
//   forward : D[128 pixels, O] = patch[128, KP] . [W | b][O, KP]^T, ReLU, bf16, TMA store into the channels-last output
//   wgrad   : D[KP, O] += patch[128, KP]^T . g[128, O]  over the CTA's pixel tiles (the SAME shared-memory image, read MN-major;
//             the gradient tile arrives by TMA), row K of D is the bias gradient; split-K over CTAs, fixed-order reduce.
// A 3-D layer is the 2-D kernel with one more tap dimension: a tile lies in one output depth slice, (n, depth) is the tile's image
// index (the channels-last output is a 4-D tensor (N * OD, OH, OW, O) for TMA), the patch holds the KD input slices around it.
namespace {

constexpr int F1_THREADS = 256;
constexpr int F1_KPMAX = 128;          // padded reduction extent (taps + ones column + zero padding), two 64-element swizzle atoms at most
constexpr int F1_ATOM = 16384;         // one 128-row x 128-byte swizzle-atom column of an operand
constexpr int F1_PATCH_MAX = 16384;    // bytes of staged fp32 input per tile

struct F1Params {
  const float* x; const float* W; const float* bias;
  const int* xrows;                    // optional: sample n of the batch is row xrows[n] of x (a shuffled epoch read straight from the dataset)
  int N, C, D, H, Wd, O, KD, KH, KW, pd, ph, pw, OD, OH, OW;
  int TW, TH, tw_sh, tiles_w, tiles_h, ntiles;
  int K, KS, NA, relu, RH, PH, PWp;    // KS = 16-wide k-steps (KP = 16 KS), NA = swizzle-atom columns of the patch operand (1 or 2)
  float* partial;                      // wgrad: [grid][KP][O]
};

// tap k -> offset of (c, t, u, v) inside the staged patch [C][KD][PH][PWp]; k == K: the ones column (-1); k > K: zero (-2)
__device__ __forceinline__ void f1_build_ktab(int* ktab, const F1Params& p) {
  for (int k = threadIdx.x; k < p.KS * 16; k += blockDim.x) {
    int off = -2;
    if (k < p.K) {
      const int T2 = p.KH * p.KW, T = p.KD * T2, c = k / T, r = k - c * T, t = r / T2, r2 = r - t * T2, u = r2 / p.KW, v = r2 - u * p.KW;
      off = ((c * p.KD + t) * p.PH + u) * p.PWp + v;
    } else if (k == p.K) off = -1;
    ktab[k] = off;
  }
}

// GEN = 0: the 2-D layers with K < 32 (two k-steps, one atom column, no depth taps) - the hot first layer of an image network, with
// everything that can be fixed at compile time fixed; GEN = 1: the general form (3-D, K < 128).
// patch element i = ((c * KD + t) * PH + r) * PWp + col  of the tile (image index nd = n * OD + od, origin (h0, w0))
template <int GEN>
__device__ __forceinline__ float f1_patch_value(const F1Params& p, int i, int n, int od, int h0, int w0) {
  const int per_t = p.PH * p.PWp;
  const int ct = i / per_t, r2 = i - ct * per_t, r = r2 / p.PWp, col = r2 - r * p.PWp;
  const int c = GEN ? ct / p.KD : ct, t = GEN ? ct - c * p.KD : 0;
  const int id = GEN ? od - p.pd + t : 0, ih = h0 - p.ph + r, iw = w0 - p.pw + col;
  float v = 0.f;
  if (n < p.N && (!GEN || (id >= 0 && id < p.D)) && ih >= 0 && ih < p.H && iw >= 0 && iw < p.Wd) {
    const size_t row = p.xrows ? (size_t)__ldg(p.xrows + n) : (size_t)n;
    v = __ldg(p.x + (((row * p.C + c) * (GEN ? p.D : 1) + id) * p.H + ih) * p.Wd + iw);
  }
  return v;
}
template <int GEN>
__device__ __forceinline__ void f1_load_patch(float* patch, const F1Params& p, int total, int n, int od, int h0, int w0) {
  for (int i = threadIdx.x; i < total; i += blockDim.x) patch[i] = f1_patch_value<GEN>(p, i, n, od, h0, w0);
}

// Software pipelining of the patch: a tile's values are fetched into registers one tile ahead (while the previous tile's MMA and
// epilogue run) and only copied to shared memory at the top of the tile - ncu had both kernels waiting on these global loads.
template <int GEN> struct F1Pre { static constexpr int N = GEN ? 8 : 6; };     // values per thread held ahead (larger patches load directly)

// Which patch elements a thread fetches never changes from tile to tile - only the tile origin does.  The element's offset relative to
// the origin and its (t, r, col) position (for the padding test) are therefore decoded ONCE per kernel; the per-tile fetch is an add and
// three range checks per value (the first version re-did four integer divisions per value and tile: ~800 instructions per thread).
template <int GEN>
__device__ __forceinline__ void f1_patch_plan(int* rel, int* trc, const F1Params& p, int total) {
  const int per_t = p.PH * p.PWp;
#pragma unroll
  for (int e = 0; e < F1Pre<GEN>::N; e++) {
    const int i = threadIdx.x + e * blockDim.x;
    rel[e] = 0; trc[e] = -1;
    if (i < total) {
      const int ct = i / per_t, r2 = i - ct * per_t, r = r2 / p.PWp, col = r2 - r * p.PWp;
      const int c = GEN ? ct / p.KD : ct, t = GEN ? ct - c * p.KD : 0;
      rel[e] = ((c * (GEN ? p.D : 1) + t) * p.H + r) * p.Wd + col;
      trc[e] = (t << 20) | (r << 10) | col;
    }
  }
}
template <int GEN>
__device__ __forceinline__ void f1_fetch_patch(float* pre, const F1Params& p, const int* rel, const int* trc, int n, int od, int h0, int w0) {
  const int d0 = GEN ? od - p.pd : 0, hh = h0 - p.ph, ww = w0 - p.pw;
  const long long row = (p.xrows && n < p.N) ? (long long)__ldg(p.xrows + n) : (long long)n;
  const float* org = p.x + ((row * p.C * (GEN ? p.D : 1) + d0) * p.H + hh) * (long long)p.Wd + ww;
#pragma unroll
  for (int e = 0; e < F1Pre<GEN>::N; e++) {
    const int t = trc[e] >> 20, r = (trc[e] >> 10) & 1023, col = trc[e] & 1023;
    const bool ok = trc[e] >= 0 && n < p.N && (!GEN || (unsigned)(d0 + t) < (unsigned)p.D) && (unsigned)(hh + r) < (unsigned)p.H &&
                    (unsigned)(ww + col) < (unsigned)p.Wd;
    pre[e] = ok ? __ldg(org + rel[e]) : 0.f;
  }
}
template <int GEN>
__device__ __forceinline__ void f1_store_patch(float* patch, const float* pre, int total) {
#pragma unroll
  for (int e = 0; e < F1Pre<GEN>::N; e++) {
    const int i = threadIdx.x + e * blockDim.x;
    if (i < total) patch[i] = pre[e];
  }
}
// tile -> image index nd = n * OD + od (the TMA coordinate), (n, od) and the tile origin inside the slice
template <int GEN>
__device__ __forceinline__ void f1_tile_origin(const F1Params& p, int tile, int& nd, int& n, int& od, int& h0, int& w0) {
  const int wt = tile % p.tiles_w, r2 = tile / p.tiles_w;
  const int ht = r2 % p.tiles_h;
  nd = r2 / p.tiles_h; w0 = wt * p.TW; h0 = ht * p.TH;
  if (GEN) { n = nd / p.OD; od = nd - n * p.OD; }
  else { n = nd; od = 0; }
}

// row m of the 128 x KP patch operand as swizzled 16-byte chunks (8 k values each): chunk j -> atom column j / 8, position j % 8;
// the CTA's threads split a row's 2 KS chunks between them (thread -> pixel m = tid % 128, chunks tid / 128, + threads / 128, ...)
template <int GEN>
__device__ __forceinline__ void f1_build_rows(uint8_t* sA, const float* patch, const int* ktab, const F1Params& p, bool pixel_ok) {
  const int m = threadIdx.x & 127, part = threadIdx.x >> 7, nparts = blockDim.x >> 7;
  const int base = (m >> p.tw_sh) * p.PWp + (m & (p.TW - 1));
  const int nchunks = GEN ? 2 * p.KS : 4;
#pragma unroll
  for (int j0 = 0; j0 < (GEN ? 16 : 4); j0++) {
    const int j = part + j0 * nparts;
    if (j >= nchunks) break;
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      float f[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int off = ktab[j * 8 + e * 2 + h];
        f[h] = off >= 0 ? patch[off + base] : (off == -1 && pixel_ok ? 1.f : 0.f);
      }
      __nv_bfloat162 t = __floats2bfloat162_rn(f[0], f[1]);
      pk[e] = *reinterpret_cast<uint32_t*>(&t);
    }
    *reinterpret_cast<uint4*>(sA + (j >> 3) * F1_ATOM + m * 128 + (((j & 7) ^ (m & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// NT = 128 (O = 64, small patches): small CTAs, five per SM - the tile loop is a chain of latencies (patch -> operand rows -> MMA ->
// TMEM -> store), and more resident CTAs hide it better than more threads per CTA (three 256-thread CTAs: 80 us vs 73 us at CIFAR
// batch 1024; ncu: "long scoreboard" on the patch)
template <int O_, int NT, int GEN>
__global__ void __launch_bounds__(NT, NT == 128 ? 5 : 2) conv_first_fprop_kernel(const __grid_constant__ CUtensorMap tmY16, const F1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                       // NA atom columns of 128 rows x 128 B
  const int NA = GEN ? p.NA : 1;
  uint8_t* sB = sA + NA * F1_ATOM;                          // NA atom columns of O_ rows x 128 B
  uint8_t* sStg = sB + NA * O_ * 128;                       // one 4 KB staging tile per warp
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(sStg + (NT / 32) * CT_STG);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 1);
  int* ktab = reinterpret_cast<int*>(mma_done + 2);
  float* patch = reinterpret_cast<float*>(ktab + F1_KPMAX);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KP = GEN ? p.KS * 16 : 32;
  constexpr int PRE = F1Pre<GEN>::N;

  if (threadIdx.x == 0) {
    mbar_init(mma_done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmY16);
  }
  if (warp == 0) tmem_alloc<O_>(tmem_slot);
  f1_build_ktab(ktab, p);
  for (int i = threadIdx.x; i < O_ * KP; i += NT) {         // [W | b | 0] as bf16, K-major, swizzled
    const int o = i / KP, k = i - o * KP;
    float v = 0.f;
    if (o < p.O) v = k < p.K ? __ldg(p.W + (size_t)o * p.K + k) : (k == p.K && p.bias ? __ldg(p.bias + o) : 0.f);
    *reinterpret_cast<__nv_bfloat16*>(sB + (k >> 6) * (O_ * 128) + o * 128 + ((((k >> 3) & 7) ^ (o & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(v);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = make_idesc(1, false, false, 128, O_);
  okl_pdl_sync();

  const int q = warp & 3;
  uint8_t* stg = sStg + warp * CT_STG;
  const uint32_t sw = (uint32_t)(lane & 7);
  uint32_t parity = 0;
  const int ptotal = p.C * p.KD * p.PH * p.PWp;
  const bool piped = ptotal <= PRE * NT;
  float pre[PRE];
  int rel[PRE], trc[PRE];
  if (piped) f1_patch_plan<GEN>(rel, trc, p, ptotal);
  if (piped && (int)blockIdx.x < p.ntiles) {
    int nd, n, od, h0, w0;
    f1_tile_origin<GEN>(p, blockIdx.x, nd, n, od, h0, w0);
    f1_fetch_patch<GEN>(pre, p, rel, trc, n, od, h0, w0);
  }
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    int nd, n, od, h0, w0;
    f1_tile_origin<GEN>(p, tile, nd, n, od, h0, w0);
    if (piped) f1_store_patch<GEN>(patch, pre, ptotal);
    else f1_load_patch<GEN>(patch, p, ptotal, n, od, h0, w0);
    __syncthreads();                                        // patch complete; every thread is past the previous tile's accumulator reads
    f1_build_rows<GEN>(sA, patch, ktab, p, true);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (piped && tile + (int)gridDim.x < p.ntiles) {        // next tile's patch: in flight during this tile's MMA and epilogue
      int nd2, n2, od2, h2, w2;
      f1_tile_origin<GEN>(p, tile + gridDim.x, nd2, n2, od2, h2, w2);
      f1_fetch_patch<GEN>(pre, p, rel, trc, n2, od2, h2, w2);
    }
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t a_lo = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16), b_lo = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
      const int nks = GEN ? p.KS : 2;
      for (int ks = 0; ks < nks; ks++)                      // k-step ks: atom column ks / 4, 32 bytes further inside its rows per step
        umma_bf16_words<0>(tmem_base, a_lo + (uint32_t)((ks >> 2) * (F1_ATOM >> 4) + (ks & 3) * 2),
                           b_lo + (uint32_t)((ks >> 2) * (O_ * 128 >> 4) + (ks & 3) * 2), TC_DESC_HI_SW128, idesc, ks ? 1u : 0u);
      umma_commit(mma_done);
    }
    mbar_wait(mma_done, parity);
    parity ^= 1;
    tc_fence_after();
    // epilogue: warp -> TMEM lane quarter q (32 pixels) and every (NT / 128)-th 64-channel chunk
    const int pbase = q * 32;
    const int bh = h0 + (pbase >> p.tw_sh);
#pragma unroll 1
    for (int ch = (warp >> 2); ch < O_ / 64; ch += NT / 128) {
      uint32_t v0[32], v1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 64);
      tmem_ld32(taddr, v0);
      tmem_ld32(taddr + 32, v1);
      tmem_ld_wait();
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t* vv = j < 4 ? v0 : v1;
        const int b = (j & 3) * 8;
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          float a = __uint_as_float(vv[b + 2 * e]), c = __uint_as_float(vv[b + 2 * e + 1]);
          if (p.relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
          __nv_bfloat162 t = __floats2bfloat162_rn(a, c);
          pk[e] = *reinterpret_cast<uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&tmY16, stg, ch * 64, w0, bh, nd);
        tma_store_commit();
      }
    }
    tc_fence_before();
  }
  if (lane == 0) tma_store_wait<0>();
  __syncthreads();
  if (warp == 0) tmem_dealloc<O_>(tmem_base);
}

template <int O_, int GEN>
__global__ void __launch_bounds__(F1_THREADS, O_ == 64 ? 3 : 2) conv_first_wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const F1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int GB = (O_ / 64) * 16384;
  const int NA = GEN ? p.NA : 1;
  const int AB = NA * F1_ATOM;
  uint8_t* sA = smem;                                       // 2 x patch operands (NA atom columns each)
  uint8_t* sG = sA + 2 * AB;                                // 2 x gradient tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + 2 * GB);
  uint64_t* g_full = bars;                                  // [2]
  uint64_t* st_free = bars + 2;                             // [2] the MMAs that read stage s are complete
  uint64_t* done = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  int* ktab = reinterpret_cast<int*>(bars + 6);
  float* patch = reinterpret_cast<float*>(ktab + F1_KPMAX);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KP = GEN ? p.KS * 16 : 32;
  constexpr int PRE = F1Pre<GEN>::N;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; s++) { mbar_init(&g_full[s], 1); mbar_init(&st_free[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmG);
  }
  if (warp == 0) tmem_alloc<O_>(tmem_slot);
  f1_build_ktab(ktab, p);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = make_idesc(1, true, true, 128, O_);
  okl_pdl_sync();

  const int ptotal = p.C * p.KD * p.PH * p.PWp;
  const bool piped = ptotal <= PRE * F1_THREADS;
  float pre[PRE];
  int rel[PRE], trc[PRE];
  if (piped) f1_patch_plan<GEN>(rel, trc, p, ptotal);
  if (piped && (int)blockIdx.x < p.ntiles) {
    int nd, n, od, h0, w0;
    f1_tile_origin<GEN>(p, blockIdx.x, nd, n, od, h0, w0);
    f1_fetch_patch<GEN>(pre, p, rel, trc, n, od, h0, w0);
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it++) {
    const int s = it & 1;
    const uint32_t use = (uint32_t)(it >> 1);               // how many times stage s has been used before
    int nd, n, od, h0, w0;
    f1_tile_origin<GEN>(p, tile, nd, n, od, h0, w0);
    if (use > 0) mbar_wait(&st_free[s], (use - 1) & 1);     // everyone: stage s (patch operand AND gradient tile) is reusable
    if (threadIdx.x == 0) {
      mbar_expect_tx(&g_full[s], GB);
#pragma unroll
      for (int j = 0; j < O_ / 64; j++) tma_load_4d(&tmG, &g_full[s], sG + s * GB + j * 16384, j * 64, w0, h0, nd);
    }
    if (piped) f1_store_patch<GEN>(patch, pre, ptotal);
    else f1_load_patch<GEN>(patch, p, ptotal, n, od, h0, w0);
    __syncthreads();
    {
      const int m = threadIdx.x & 127;
      const bool ok = (w0 + (m & (p.TW - 1))) < p.OW && (h0 + (m >> p.tw_sh)) < p.OH && n < p.N;
      f1_build_rows<GEN>(sA + s * AB, patch, ktab, p, ok);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (piped && tile + (int)gridDim.x < p.ntiles) {
      int nd2, n2, od2, h2, w2;
      f1_tile_origin<GEN>(p, tile + gridDim.x, nd2, n2, od2, h2, w2);
      f1_fetch_patch<GEN>(pre, p, rel, trc, n2, od2, h2, w2);
    }
    if (threadIdx.x == 0) {
      mbar_wait(&g_full[s], use & 1);
      tc_fence_after();
      // A = patch^T, MN-major: 16 pixel rows per MMA, M = the KP k values (64 per atom column, LBO = column distance); with one
      // atom column M rows 64..127 alias an in-bounds region and their results are never read
      const uint32_t a_lo = ((smem_u32(sA + s * AB) & 0x3FFFFu) >> 4) | ((uint32_t)((NA > 1 ? F1_ATOM : 1024) >> 4) << 16);
      const uint32_t g_lo = ((smem_u32(sG + s * GB) & 0x3FFFFu) >> 4) | ((16384u >> 4) << 16);
#pragma unroll
      for (int kk = 0; kk < 8; kk++)
        umma_bf16_words<0>(tmem_base, a_lo + (uint32_t)(kk * 128), g_lo + (uint32_t)(kk * 128), TC_DESC_HI_SW128, idesc, (it | kk) ? 1u : 0u);
      umma_commit(&st_free[s]);
    }
  }
  if (threadIdx.x == 0) umma_commit(done);
  mbar_wait(done, 0);
  tc_fence_after();
  if (warp < 4 && warp * 32 < KP) {                         // TMEM lanes = k rows: warp w reads lanes 32 w .. 32 w + 31
    float* dst = p.partial + ((size_t)blockIdx.x * KP + warp * 32 + lane) * O_;
    const bool any = (int)blockIdx.x < p.ntiles;
#pragma unroll 1
    for (int c = 0; c < O_; c += 32) {
      uint32_t v[32];
      if (any) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = 0u;
      }
      if (warp * 32 + lane < KP) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(dst + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<O_>(tmem_base);
}

// one warp per output element: lanes stride over the CTAs' partials, fixed-order shuffle tree
__global__ void conv_first_wgrad_reduce(const float* __restrict__ partial, float* __restrict__ dW, float* __restrict__ db, int O, int Opad, int K, int KP,
                                        int nblk) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (idx >= (K + 1) * O) return;
  const int k = idx / O, o = idx - k * O;
  float s0 = 0.f, s1 = 0.f;
  int b = lane;
  for (; b + 32 < nblk; b += 64) {
    s0 += partial[((size_t)b * KP + k) * Opad + o];
    s1 += partial[((size_t)(b + 32) * KP + k) * Opad + o];
  }
  if (b < nblk) s0 += partial[((size_t)b * KP + k) * Opad + o];
  float s = s0 + s1;
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
  if (lane == 0) {
    if (k < K) dW[(size_t)o * K + k] = s;
    else if (db) db[o] = s;
  }
}

bool f1_plan(okl_ctx* ctx, const okl_conv_desc* d, F1Params& p, int& opad) {
  if (ctx->cc < 100 || ct_env("OKL_CONV_FIRST_TC", 1) == 0) return false;
  if ((d->nd != 2 && d->nd != 3) || d->x_layout != OKL_NCX || d->y_layout != OKL_NXC) return false;
  for (int i = 0; i < d->nd; i++)
    if (d->s[i] != 1) return false;
  memset(&p, 0, sizeof(p));
  const int o3 = d->nd == 3 ? 1 : 0;                       // spatial index of H
  p.N = d->N; p.C = d->C; p.O = d->O;
  p.D = o3 ? d->in[0] : 1; p.KD = o3 ? d->k[0] : 1; p.pd = o3 ? d->p[0] : 0;
  p.H = d->in[o3]; p.Wd = d->in[o3 + 1]; p.KH = d->k[o3]; p.KW = d->k[o3 + 1]; p.ph = d->p[o3]; p.pw = d->p[o3 + 1];
  p.K = p.C * p.KD * p.KH * p.KW;
  if (p.K + 1 > F1_KPMAX || p.O % 64 || p.O > 256 || p.N < 1) return false;
  p.KS = (p.K + 1 + 15) / 16; if (p.KS < 2) p.KS = 2;
  p.NA = p.KS > 4 ? 2 : 1;
  opad = p.O <= 64 ? 64 : (p.O <= 128 ? 128 : 256);
  if (opad != p.O) return false;
  p.OD = p.D + 2 * p.pd - p.KD + 1; p.OH = p.H + 2 * p.ph - p.KH + 1; p.OW = p.Wd + 2 * p.pw - p.KW + 1;
  if (p.OD < 1 || p.OH < 1 || p.OW < 1 || (long long)p.N * p.OD > 0x7fffffffLL) return false;
  // pixel tile TW x (128 / TW) inside one slice: the candidate that covers the slice with the fewest out-of-range pixels
  long long best = -1;
  for (int tw = 32; tw >= 8; tw >>= 1) {
    if (tw > 8 && tw / 2 >= p.OW) continue;                 // narrower images: the next candidate fits them without waste
    const int th = 128 / tw;
    const long long cover = (long long)((p.OW + tw - 1) / tw) * tw * ((p.OH + th - 1) / th) * th;
    if ((p.C * p.KD * (th + p.KH - 1) * (tw + p.KW - 1)) * 4 > F1_PATCH_MAX) continue;
    if (best < 0 || cover < best) { best = cover; p.TW = tw; }
  }
  if (best < 0) return false;
  p.TH = 128 / p.TW;
  p.tw_sh = ct_ilog2(p.TW);
  p.tiles_w = (p.OW + p.TW - 1) / p.TW; p.tiles_h = (p.OH + p.TH - 1) / p.TH;
  const long long nt = (long long)p.tiles_w * p.tiles_h * p.N * p.OD;
  if (nt > 0x7fffffffLL) return false;
  p.ntiles = (int)nt;
  p.RH = 32 / p.TW;
  p.PH = p.TH + p.KH - 1; p.PWp = p.TW + p.KW - 1;
  return true;
}

}  // namespace

extern "C" int okl_conv_first_supported(okl_ctx* ctx, const okl_conv_desc* d) {
  if (!ctx || !d) return 0;
  F1Params p;
  int opad;
  return f1_plan(ctx, d, p, opad) ? 1 : 0;
}

// y16 (N, *out, O) bf16 = act(conv(x (N, C, *in) fp32, W) + b); d->act NONE or RELU
extern "C" int okl_conv_first_fprop(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* x_rows_i32, const void* W, const void* bias, void* y16) {
  OKL_REQUIRE(ctx && d && x && W && y16, "NULL argument");
  F1Params p;
  int opad;
  OKL_REQUIRE(f1_plan(ctx, d, p, opad), "okl_conv_first_fprop: unsupported shape / layout");
  OKL_REQUIRE(d->act == OKL_ACT_NONE || d->act == OKL_ACT_RELU, "okl_conv_first_fprop: only ReLU can be fused");
  p.x = (const float*)x; p.xrows = (const int*)x_rows_i32; p.W = (const float*)W; p.bias = (const float*)bias; p.relu = d->act == OKL_ACT_RELU;
  CUtensorMap ty;
  int s = ct_map_nhwc(&ty, y16, 2, p.O, p.OW, p.OH, p.N * p.OD, 64, p.TW, p.RH, 1);
  if (s != OKL_OK) return s;
  const int ptotal = p.C * p.KD * p.PH * p.PWp;
  const int gen = (p.KS > 2 || p.KD > 1 || p.D > 1) ? 1 : 0;
  const int nt = (opad == 64 && ptotal <= (gen ? 8 : 6) * 128 && p.NA == 1) ? 128 : 256;
  const int smem = p.NA * F1_ATOM + p.NA * opad * 128 + (nt / 32) * CT_STG + 64 + F1_KPMAX * 4 + ptotal * 4 + 1024;
  int per_sm = nt == 128 ? 5 : 2;
  if (per_sm * (smem + 1024) > 227 * 1024) per_sm = (227 * 1024) / (smem + 1024);
  if (per_sm < 1) per_sm = 1;
  const int grid = p.ntiles < per_sm * ctx->sm_count ? p.ntiles : per_sm * ctx->sm_count;
#define F1_FP(OO, NT_, G_)                                                                                            \
  {                                                                                                                   \
    static int set = 0;                                                                                               \
    if (set < smem) { OKL_CHECK_CUDA(cudaFuncSetAttribute(conv_first_fprop_kernel<OO, NT_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); set = 160 * 1024; } \
    okl_launch(ctx, conv_first_fprop_kernel<OO, NT_, G_>, dim3(grid), dim3(NT_), smem, ty, p);                         \
  }
  if (gen) {
    if (opad == 64 && nt == 128) F1_FP(64, 128, 1) else if (opad == 64) F1_FP(64, 256, 1) else if (opad == 128) F1_FP(128, 256, 1) else F1_FP(256, 256, 1)
  } else {
    if (opad == 64 && nt == 128) F1_FP(64, 128, 0) else if (opad == 64) F1_FP(64, 256, 0) else if (opad == 128) F1_FP(128, 256, 0) else F1_FP(256, 256, 0)
  }
#undef F1_FP
  OKL_LAUNCHED(ctx, "conv_first_fprop");
  return OKL_OK;
}

// dW (O, C, *k) fp32 and db (O) fp32 (optional) from x (N, C, *in) fp32 and g16 (N, *out, O) bf16
extern "C" int okl_conv_first_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* x_rows_i32, const void* g16, void* dW, void* db) {
  OKL_REQUIRE(ctx && d && x && g16 && dW, "NULL argument");
  F1Params p;
  int opad;
  OKL_REQUIRE(f1_plan(ctx, d, p, opad), "okl_conv_first_wgrad: unsupported shape / layout");
  p.x = (const float*)x; p.xrows = (const int*)x_rows_i32;
  CUtensorMap tg;
  int s = ct_map_nhwc(&tg, g16, 2, p.O, p.OW, p.OH, p.N * p.OD, 64, p.TW, p.TH, 1);
  if (s != OKL_OK) return s;
  const int KP = p.KS * 16;
  const int smem = 2 * p.NA * F1_ATOM + 2 * (opad / 64) * 16384 + 64 + F1_KPMAX * 4 + p.C * p.KD * p.PH * p.PWp * 4 + 1024;
  int per_sm = opad == 64 ? 3 : 2;
  if (per_sm * (smem + 1024) > 227 * 1024) per_sm = (227 * 1024) / (smem + 1024);
  if (per_sm < 1) per_sm = 1;
  const int grid = p.ntiles < per_sm * ctx->sm_count ? p.ntiles : per_sm * ctx->sm_count;
  s = okl_ensure_workspace(ctx, (size_t)grid * KP * opad * sizeof(float));
  if (s != OKL_OK) return s;
  p.partial = (float*)ctx->workspace;
  const int gen = (p.KS > 2 || p.KD > 1 || p.D > 1) ? 1 : 0;
#define F1_WG(OO, G_)                                                                                                 \
  {                                                                                                                   \
    static int set = 0;                                                                                               \
    if (set < smem) { OKL_CHECK_CUDA(cudaFuncSetAttribute(conv_first_wgrad_kernel<OO, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); set = 200 * 1024; } \
    okl_launch(ctx, conv_first_wgrad_kernel<OO, G_>, dim3(grid), dim3(F1_THREADS), smem, tg, p);                       \
  }
  if (gen) {
    if (opad == 64) F1_WG(64, 1) else if (opad == 128) F1_WG(128, 1) else F1_WG(256, 1)
  } else {
    if (opad == 64) F1_WG(64, 0) else if (opad == 128) F1_WG(128, 0) else F1_WG(256, 0)
  }
#undef F1_WG
  OKL_LAUNCHED(ctx, "conv_first_wgrad");
  const int total = (p.K + 1) * p.O;
  conv_first_wgrad_reduce<<<(total * 32 + 255) / 256, 256, 0, ctx->stream>>>(p.partial, (float*)dW, (float*)db, p.O, opad, p.K, KP, grid);
  OKL_LAUNCHED(ctx, "conv_first_wgrad_reduce");
  return OKL_OK;
}

// ================================================================================================ first layer, horizontal taps pre-expanded
// Conv2DLayer(3, 64, 3, 1, 1) on the input images (okl.py:790-813 / 869-871) once more.  The small-K kernels above build the im2col
// operand of every pixel tile in software (~600 instructions per pixel and tile: 61 + 70 us at batch 1024, issue-bound at 0.32-0.37 of
// HBM).  With C . KW <= 16 the horizontal taps fit the channel dimension instead: ONE memory-bound pass writes
//     xe[n, h, ow, c * KW + v] = x[n, c, h, ow + v - pw]        (bf16, channels-last, 16 channels of which C . KW are real)
// and the layer becomes an ordinary channels-last convolution with 16 input channels, a KH x 1 filter and padding (ph, 0) - the
// TMA-fed implicit GEMM (conv_tc_kernel: one 64-channel box per pixel, zero-filled past channel 16) and the gradient-slab weight
// gradient take it as they take any layer with a partial channel chunk; bias gradient from the ones slot.  Nothing is gathered per tap.
// filters: w2e[o][u][c * KW + v] = bf16(W[o][c][u][v]); dW[o][c][u][v] = dWe[o][c * KW + v][u].
namespace {

constexpr int VX_C = 16;

__global__ void vexp_expand_kernel(const float* __restrict__ x, const int* __restrict__ rows, uint4* __restrict__ xe, int N, int C, int H, int W, int OW, int KW,
                                   int pw) {
  // one thread per output pixel: 16 bf16 = two 16-byte stores; the reads of a warp are KW . C shifted, coalesced row segments
  const size_t total = (size_t)N * H * OW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ow = (int)(i % OW);
    const size_t r = i / OW;
    const int h = (int)(r % H), n = (int)(r / H);
    const size_t row = rows ? (size_t)__ldg(rows + n) : (size_t)n;
    const float* src = x + (row * C * H + h) * (size_t)W;
    float f[VX_C];
#pragma unroll
    for (int k = 0; k < VX_C; k++) f[k] = 0.f;
    for (int c = 0; c < C; c++)
      for (int v = 0; v < KW; v++) {
        const int iw = ow + v - pw;
        if (iw >= 0 && iw < W) f[c * KW + v] = __ldg(src + (size_t)c * H * W + iw);
      }
    uint32_t pk[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
      pk[k] = *reinterpret_cast<uint32_t*>(&t);
    }
    xe[2 * i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    xe[2 * i + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

__global__ void vexp_bank_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ w2e, int O, int C, int KH, int KW) {
  const int total = O * KH * VX_C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ce = i % VX_C, u = (i / VX_C) % KH, o = i / (VX_C * KH);
    const int c = ce / KW, v = ce - c * KW;
    w2e[i] = __float2bfloat16(c < C ? W[(((size_t)o * C + c) * KH + u) * KW + v] : 0.f);
  }
}

__global__ void vexp_dw_kernel(const float* __restrict__ dWe, float* __restrict__ dW, int O, int C, int KH, int KW) {
  const int total = O * C * KH * KW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % KW, u = (i / KW) % KH, c = (i / (KW * KH)) % C, o = i / (KW * KH * C);
    dW[i] = dWe[((size_t)o * VX_C + c * KW + v) * KH + u];
  }
}

// the channels-last convolution the layer turns into
bool vexp_desc(const okl_conv_desc* d, okl_conv_desc& e) {
  if (d->nd != 2 || d->s[0] != 1 || d->s[1] != 1 || d->x_layout != OKL_NCX || d->y_layout != OKL_NXC) return false;
  if (d->C < 1 || d->C * d->k[1] > VX_C || d->k[0] < 2 || d->k[0] > 4 || d->k[1] < 1 || d->p[1] > d->k[1] - 1 || d->N < 1) return false;
  const int OW = d->in[1] + 2 * d->p[1] - d->k[1] + 1;
  if (OW < 1) return false;
  e = *d;
  e.C = VX_C;
  e.in[1] = OW; e.k[1] = 1; e.p[1] = 0;
  e.x_layout = OKL_NXC;
  return true;
}

}  // namespace

extern "C" int okl_conv_vexp_supported(okl_ctx* ctx, const okl_conv_desc* d) {
  if (!ctx || !d || ct_env("OKL_CONV_VEXP", 1) == 0) return 0;
  okl_conv_desc e;
  if (!vexp_desc(d, e) || !ct_supported(ctx, &e)) return 0;
  GsPlan gp;
  return (gs_plan(ctx, &e, true, gp) && gp.db_fused) ? 1 : 0;
}

// the derived channels-last descriptor (16 input channels, KH x 1 filter) for okl_conv2_fprop / okl_conv2_wgrad
extern "C" int okl_conv_vexp_desc(okl_ctx* ctx, const okl_conv_desc* d, okl_conv_desc* out) {
  OKL_REQUIRE(ctx && d && out, "NULL argument");
  OKL_REQUIRE(vexp_desc(d, *out), "okl_conv_vexp_desc: unsupported shape / layout");
  return OKL_OK;
}

// xe16 (N, H, OW, 16) bf16 from x (N, C, H, W) fp32 (optionally rows x_rows_i32[n] of a dataset)
extern "C" int okl_conv_vexp_expand(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* x_rows_i32, void* xe16) {
  OKL_REQUIRE(ctx && d && x && xe16, "NULL argument");
  okl_conv_desc e;
  OKL_REQUIRE(vexp_desc(d, e), "okl_conv_vexp_expand: unsupported shape / layout");
  const size_t total = (size_t)d->N * d->in[0] * e.in[1];
  vexp_expand_kernel<<<okl_grid_1d(ctx, total, 256, 16), 256, 0, ctx->stream>>>((const float*)x, (const int*)x_rows_i32, (uint4*)xe16, d->N, d->C, d->in[0], d->in[1],
                                                                               e.in[1], d->k[1], d->p[1]);
  OKL_LAUNCHED(ctx, "conv_vexp_expand");
  return OKL_OK;
}

// w2e16 (O, KH, 16) bf16: the K-major filter bank of the derived convolution
extern "C" int okl_conv_vexp_bank(okl_ctx* ctx, const okl_conv_desc* d, const void* W, void* w2e16) {
  OKL_REQUIRE(ctx && d && W && w2e16, "NULL argument");
  okl_conv_desc e;
  OKL_REQUIRE(vexp_desc(d, e), "okl_conv_vexp_bank: unsupported shape / layout");
  vexp_bank_kernel<<<okl_grid_1d(ctx, (size_t)d->O * d->k[0] * VX_C, 256), 256, 0, ctx->stream>>>((const float*)W, (__nv_bfloat16*)w2e16, d->O, d->C, d->k[0], d->k[1]);
  OKL_LAUNCHED(ctx, "conv_vexp_bank");
  return OKL_OK;
}

// dW (O, C, KH, KW) fp32 from the derived convolution's filter gradient dWe (O, 16, KH, 1)
extern "C" int okl_conv_vexp_dw(okl_ctx* ctx, const okl_conv_desc* d, const void* dWe, void* dW) {
  OKL_REQUIRE(ctx && d && dWe && dW, "NULL argument");
  okl_conv_desc e;
  OKL_REQUIRE(vexp_desc(d, e), "okl_conv_vexp_dw: unsupported shape / layout");
  vexp_dw_kernel<<<okl_grid_1d(ctx, (size_t)d->O * d->C * d->k[0] * d->k[1], 256), 256, 0, ctx->stream>>>((const float*)dWe, (float*)dW, d->O, d->C, d->k[0], d->k[1]);
  OKL_LAUNCHED(ctx, "conv_vexp_dw");
  return OKL_OK;
}
