// CoAM position attention on the sm_100a tensor cores, every head dim C = 16 .. 384 (multiple of 16): all branches of
// W32 / W48 at every CoAM placement.  S = QK^T and O stay in TMEM and both GEMMs are tcgen05.mma; channels are handled
// in 64-wide blocks (one SWIZZLE_128B block = [128 rows][64 channels]): Q is resident as ceil(C/64) blocks, the K and V
// blocks of a key tile stream through one ring (K_0 .. K_n-1, V_0 .. V_n-1), S accumulates over the K blocks and V block i
// feeds O columns [64 i, 64 i + 64).  TMEM: S 128 columns + O C columns (<= 512).
//
//   per CTA: 128 queries (one thread of the 4 soft-max warps per query row = TMEM lane), loop over 128-key tiles:
//     warp 0  TMA: Q once, K/V tiles through a 2-stage ring (SWIZZLE_128B, 64-channel boxes, zero fill above C)
//     warp 1  MMA: S[128x128] = Q K^T (K-major A and B), then O[128xC] += P V  (P from shared memory, V as an
//             MN-major B operand: the tile is [keys][channels], exactly as it lies in the act16 tensor)
//     warps 2..5  soft-max: two passes over the S row in TMEM (max, then exp2 -> 16-bit P into shared memory),
//             running max / sum in registers, O rescaled in TMEM only when a row maximum grows
//   C <= 64: two CTAs per SM interleave, so one CTA's soft-max overlaps the other's MMAs (C > 64: one CTA per SM; those
//   branches are 12 % of the attention FLOPs).
#include "common.cuh"
#include "internal.h"

namespace buctd {

#ifdef BUCTD_ROLE_TIMERS
__device__ unsigned long long g_pa_cycles[16];
#define PT_DECL(n) long long n = 0
#define PT_BEGIN(t) const long long t = clock64()
#define PT_ADD(a, t) a += clock64() - t
#else
#define PT_DECL(n)
#define PT_BEGIN(t)
#define PT_ADD(a, t)
#endif

constexpr int kTcThreads = 192;
constexpr int kTcTile = 128;
constexpr int kTcBlk = 128 * 128;      // one [128 rows][128 B] SWIZZLE_128B block

struct PosAttTcParams {
  int nrows, W, H, c, n2;              // rows per crop (padded), geometry, channels, O columns in TMEM (C, + the ones column)
  int nkc, ns;                         // 64-channel blocks; K/V ring slots
  int n_last;                          // MMA N of the last V block (its channels, + the ones column, rounded up to 16)
  int ones;                            // row sums come from a column of ones planted in the last V block
  int s_first;                         // two-tile kernel: issue S(j+1) before PV(j)
  int dbg;                             // BUCTD_ROLE_TIMERS builds: timing-only switches (BUCTD_PA_DBG)
  int tmem_cols;
  float drop_p;                        // training: dropout of the probabilities (hash mask), 0: none
  uint32_t drop_seed;
  float* lse2;                         // optional: log2-sum-exp per query row
  const uint32_t* seed_xor;            // optional device word XOR-ed into drop_seed
  int o_cs;
  float sc;                            // 1/sqrt(C) * log2(e)
  void* o;
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major, SWIZZLE_128B descriptor: the tile is [K rows][64 MN elements = 128 B]; 8-row groups are 1024 B apart
// (SBO); LBO would be the distance between 64-element column blocks (not used for N <= 64)
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(kTcBlk >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// non-saturating pack (the values are probabilities scaled by at most 2^8)
template <typename T> __device__ __forceinline__ uint32_t pack_ns(float a, float b);
template <> __device__ __forceinline__ uint32_t pack_ns<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack_ns<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <typename T> __device__ __forceinline__ uint32_t ex2_pack(float a, float b);
template <> __device__ __forceinline__ uint32_t ex2_pack<__half>(float a, float b) {
  uint32_t h = T16<__half>::pack2(a, b), r;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(h));
  return r;
}
template <> __device__ __forceinline__ uint32_t ex2_pack<__nv_bfloat16>(float a, float b) {
  uint32_t h = T16<__nv_bfloat16>::pack2(a, b), r;
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(h));
  return r;
}

// -DBUCTD_PA_PACKED_EXP: ex2.approx.f16x2 in the fp16 fast path (one MUFU op per two scores).  Measured on B200: 13.44 ms
// per 64-crop forward against 13.34 ms with two float32 ex2 + a pack -- the soft-max warps are not MUFU-throughput bound
// (DESIGN.md 3.3); off by default.
#ifdef BUCTD_PA_PACKED_EXP
template <typename T> constexpr bool kPackedExp = false;
template <> constexpr bool kPackedExp<__half> = true;
#else
template <typename T> constexpr bool kPackedExp = false;
#endif

// one group of four soft-max warps working on one 128-query tile (thread = query row = TMEM lane)
struct PaGroup {
  uint8_t* s_p;                        // this group's P tile [2 blocks][128][128 B]
  float* s_bias;                       // [2][128] additive key mask (double buffered)
  uint32_t* s_mask;                    // [2][4]
  uint64_t *s_full, *p_full, *o_done;  // this group's S-ready / P-ready / PV-done barriers
  uint64_t* o_half;                    // first half of the PV MMAs done: keys 0..63 of the P tile may be overwritten
  uint64_t* kv_full;                   // the K/V ring
  uint8_t* s_kv;
  uint32_t tm_s, tm_o;                 // this group's S and O accumulators
  int m0;                              // first query row of the tile
  int etid;                            // 0..127 inside the group
  int bar_id;                          // named barrier of the group
  bool plant;                          // this group plants the ones column into the V tiles
};

template <typename T, bool TRAIN>
__device__ __forceinline__ void pa_softmax_role(const PosAttTcParams& p, const PaGroup& g, int b, int row_base, int ntiles, int NKC, int NS) {
  // ---- soft-max warps: thread = query row --------------------------------------------------------------
  // Single pass per tile in the common case: P = exp2(s*sc + bias - m_run) with the running maximum of EARLIER
  // tiles; the tile's own maximum is tracked on the side and only when it exceeds m_run by more than 8 (P would
  // leave the safe range of the 16-bit type) is the row state updated (O rescaled in TMEM) and the tile redone.
  // The result is independent of the reference maximum, so this is exact.  Row sums: when a channel slot is free
  // (C < 64) a column of ones is planted in the V tile, so O[:, C] accumulates sum_k P with the tensor cores.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int etid = g.etid;
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  const int Wp = p.W + 1;
  const bool ones = p.ones != 0;               // a free channel slot exists for the ones column
  const uint32_t drop_thr = (TRAIN && p.drop_p > 0.f) ? drop_threshold(p.drop_p) : 0u;
  const uint32_t drop_seed = TRAIN ? (p.drop_seed ^ ((p.drop_p > 0.f && p.seed_xor) ? *p.seed_xor : 0u)) : 0u;
  float m_run = -INFINITY, l_run = 0.f;
  uint8_t* const s_p = g.s_p; float* const s_bias = g.s_bias;
  uint64_t* const s_full = g.s_full; uint64_t* const p_full = g.p_full; uint64_t* const o_done = g.o_done;
  uint64_t* const o_half = g.o_half;
  uint64_t* const kv_full = g.kv_full; uint8_t* const s_kv = g.s_kv;
  const uint32_t tm_s = g.tm_s, tm_o = g.tm_o;
  const int m0 = g.m0;
  uint8_t* prow = s_p + row * 128;
  const int sw = row & 7;
  PT_DECL(x_bar); PT_DECL(x_s); PT_DECL(x_o); PT_DECL(x_main); PT_DECL(x_v); PT_DECL(x_tot); PT_DECL(x_redo); PT_DECL(x_ld); PT_DECL(x_ldi); PT_DECL(x_ch);
  PT_BEGIN(x_start);
  // key r of a crop is a pad iff r % (W + 1) == W (pad column) or r >= H (W + 1) (bottom pad row, rows beyond the crop);
  // the two remainders are carried from tile to tile (no division per tile)
  const int n_valid = p.H * Wp;
  const int step = kTcTile % Wp;
  int rem_e = etid % Wp;                       // (j * 128 + etid) % Wp: this thread's key of the tile
  int rem_0 = 0;                               // (j * 128) % Wp
  for (int j = 0; j < ntiles; ++j) {
    float* bias = s_bias + (j & 1) * kTcTile;
    bias[etid] = (j * kTcTile + etid < n_valid && rem_e != p.W) ? 0.f : -INFINITY;
    const int c_pad = rem_0 <= p.W ? p.W - rem_0 : p.W - rem_0 + Wp;      // first pad column of the tile
    rem_e += step; if (rem_e >= Wp) rem_e -= Wp;
    rem_0 += step; if (rem_0 >= Wp) rem_0 -= Wp;
    { PT_BEGIN(t0); asm volatile("bar.sync %0, 128;" ::"r"(g.bar_id) : "memory"); PT_ADD(x_bar, t0); }
    { PT_BEGIN(t0); mbar_wait(s_full, j & 1); PT_ADD(x_s, t0); }
    tc_fence_after();
    bool first = true;
    bool p_hi_busy = false;                                 // keys 64..127 of the P tile may still be read by the previous PV
    float ls = 0.f;
    for (;;) {
      // on the very first tile (or on a redo) the row maximum must be known before P can be formed
      if (m_run == -INFINITY || !first) {
        float mx = -INFINITY;
#pragma unroll 1
        for (int c0 = 0; c0 < kTcTile; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tm_s + lane_off + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bq = *reinterpret_cast<const float4*>(bias + c0 + i);
            mx = fmaxf(mx, fmaf(__uint_as_float(v[i]), p.sc, bq.x));
            mx = fmaxf(mx, fmaf(__uint_as_float(v[i + 1]), p.sc, bq.y));
            mx = fmaxf(mx, fmaf(__uint_as_float(v[i + 2]), p.sc, bq.z));
            mx = fmaxf(mx, fmaf(__uint_as_float(v[i + 3]), p.sc, bq.w));
          }
        }
        if (mx == -INFINITY && j + 1 < ntiles) mx = -1e30f;      // a tile of masked keys only: keep the state finite
        const float m_new = fmaxf(m_run, mx);
        const float m_use = m_new == -INFINITY ? 0.f : m_new;
        const float corr = exp2f(m_run - m_use);            // 0 when m_run = -inf
        if (j > 0) {                                        // rescale O (and l, which may live in O[:, C])
          mbar_wait(o_done, (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffff, corr != 1.f)) {
#pragma unroll 1
            for (int c0 = 0; c0 < p.n2; c0 += 16) {
              uint32_t v[16];
              tmem_ld16(tm_o + lane_off + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * corr);
              tmem_st16(tm_o + lane_off + c0, v);
            }
            tmem_st_wait();
          }
        }
        l_run *= corr;
        m_run = m_new;
      } else if (j > 0) {
        PT_BEGIN(t0);
        mbar_wait(o_half, (j - 1) & 1);                     // keys 0..63 of the P tile are free once the first four PV MMAs are done
        PT_ADD(x_o, t0);
        tc_fence_after();
        p_hi_busy = true;
      }
      PT_BEGIN(t_main);
      const float m_use = m_run == -INFINITY ? 0.f : m_run;
      float mx2 = -INFINITY;
      ls = 0.f;
      // one 32-column chunk: scores -> P (16 bit) -> this row of the P tile
      auto chunk = [&](int c0, const uint32_t (&v)[32]) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float4 bq = *reinterpret_cast<const float4*>(bias + c0 + 2 * i);
          const float t0 = fmaf(__uint_as_float(v[2 * i]), p.sc, bq.x), t1 = fmaf(__uint_as_float(v[2 * i + 1]), p.sc, bq.y);
          const float t2 = fmaf(__uint_as_float(v[2 * i + 2]), p.sc, bq.z), t3 = fmaf(__uint_as_float(v[2 * i + 3]), p.sc, bq.w);
          mx2 = fmaxf(fmaxf(mx2, fmaxf(t0, t1)), fmaxf(t2, t3));
          pk[i] = ex2_pack<T>(t0 - m_use, t1 - m_use);
          pk[i + 1] = ex2_pack<T>(t2 - m_use, t3 - m_use);
          if (!ones) {
            const float2 f0 = T16<T>::unpack2(pk[i]), f1 = T16<T>::unpack2(pk[i + 1]);
            ls += (f0.x + f0.y) + (f1.x + f1.y);
          }
          if (TRAIN && drop_thr != 0u) {           // the row sum above is over the UNDROPPED probabilities; 1 / (1 - p) goes into `inv`
            const uint32_t kb = (uint32_t)(j * kTcTile + c0 + 2 * i);
            const uint32_t k0 = drop_hash(drop_seed, (uint32_t)b, (uint32_t)(m0 + row), kb) >= drop_thr ? 0x0000FFFFu : 0u;
            const uint32_t k1 = drop_hash(drop_seed, (uint32_t)b, (uint32_t)(m0 + row), kb + 1) >= drop_thr ? 0xFFFF0000u : 0u;
            const uint32_t k2 = drop_hash(drop_seed, (uint32_t)b, (uint32_t)(m0 + row), kb + 2) >= drop_thr ? 0x0000FFFFu : 0u;
            const uint32_t k3 = drop_hash(drop_seed, (uint32_t)b, (uint32_t)(m0 + row), kb + 3) >= drop_thr ? 0xFFFF0000u : 0u;
            pk[i] &= (k0 | k1);
            pk[i + 1] &= (k2 | k3);
          }
        }
        uint8_t* blk = prow + (c0 >> 6) * kTcBlk;
        const int ci = (c0 & 63) >> 3;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          *reinterpret_cast<uint4*>(blk + (((ci + ch) ^ sw) << 4)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
      };
      // fast variant (row sums come from the ones column): no per-element mask, 3.5 instructions per score --
      // raw max, one FFMA, one ex2, half a pack; masked keys are zeroed afterwards in the P tile
      float rawmax = -INFINITY;
      const float neg_m = -m_use;
      auto chunk_fast = [&](int c0, const uint32_t (&v)[32]) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float s0 = __uint_as_float(v[2 * i]), s1 = __uint_as_float(v[2 * i + 1]);
          rawmax = fmaxf(rawmax, fmaxf(s0, s1));
#ifdef BUCTD_ROLE_TIMERS
          if (p.dbg & 1) { pk[i] = pack_ns<T>(fmaf(s0, p.sc, neg_m), fmaf(s1, p.sc, neg_m)); continue; }      // timing only: no exponentials
#endif
          if (kPackedExp<T>) pk[i] = ex2_pack<T>(fmaf(s0, p.sc, neg_m), fmaf(s1, p.sc, neg_m));
          else pk[i] = pack_ns<T>(ex2f(fmaf(s0, p.sc, neg_m)), ex2f(fmaf(s1, p.sc, neg_m)));
        }
        uint8_t* blk = prow + (c0 >> 6) * kTcBlk;
        const int ci = (c0 & 63) >> 3;
#ifdef BUCTD_ROLE_TIMERS
        if (p.dbg & 2) {                                   // timing only: one store instead of four
          *reinterpret_cast<uint4*>(blk + ((ci ^ sw) << 4)) = make_uint4(pk[0] ^ pk[4] ^ pk[8] ^ pk[12], pk[1] ^ pk[5] ^ pk[9] ^ pk[13], pk[2] ^ pk[6] ^ pk[10] ^ pk[14], pk[3] ^ pk[7] ^ pk[11] ^ pk[15]);
          return;
        }
#endif
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          *reinterpret_cast<uint4*>(blk + (((ci + ch) ^ sw) << 4)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
      };
      // TMEM reads are software pipelined: the load of chunk c+1 is in flight while chunk c is processed
      {
        uint32_t va[32], vb[32];
#ifdef BUCTD_ROLE_TIMERS
#define PA_LD(c, r) do { if (!(p.dbg & 4)) tmem_ld32(tm_s + lane_off + (c), r); else { _Pragma("unroll") for (int i_ = 0; i_ < 32; ++i_) r[i_] = __float_as_uint((float)(int)((threadIdx.x ^ (unsigned)j) + i_ + (c)) * 1e-4f); } } while (0)
#else
#define PA_LD(c, r) tmem_ld32(tm_s + lane_off + (c), r)
#endif
        { PT_BEGIN(t0); PA_LD(0, va); PT_ADD(x_ldi, t0); }
        { PT_BEGIN(t0); tmem_ld_wait(); PT_ADD(x_ld, t0); }
        { PT_BEGIN(t0); PA_LD(32, vb); PT_ADD(x_ldi, t0); }
        { PT_BEGIN(t0); if (ones) chunk_fast(0, va); else chunk(0, va); PT_ADD(x_ch, t0); }
        { PT_BEGIN(t0); tmem_ld_wait(); PT_ADD(x_ld, t0); }
        { PT_BEGIN(t0); PA_LD(64, va); PT_ADD(x_ldi, t0); }
        { PT_BEGIN(t0); if (ones) chunk_fast(32, vb); else chunk(32, vb); PT_ADD(x_ch, t0); }
        { PT_BEGIN(t0); tmem_ld_wait(); PT_ADD(x_ld, t0); }
        { PT_BEGIN(t0); PA_LD(96, vb); PT_ADD(x_ldi, t0); }
        if (p_hi_busy) {
          PT_BEGIN(t0);
          mbar_wait(o_done, (j - 1) & 1);
          PT_ADD(x_o, t0);
          tc_fence_after();
          p_hi_busy = false;
        }
        { PT_BEGIN(t0); if (ones) chunk_fast(64, va); else chunk(64, va); PT_ADD(x_ch, t0); }
        { PT_BEGIN(t0); tmem_ld_wait(); PT_ADD(x_ld, t0); }
        { PT_BEGIN(t0); if (ones) chunk_fast(96, vb); else chunk(96, vb); PT_ADD(x_ch, t0); }
      }
      if (ones) {
        mx2 = rawmax * p.sc;
        // zero the masked keys of this row's P: the pad columns, and everything from the bottom pad row on
        const int c_end = min(kTcTile, n_valid - j * kTcTile);           // keys [c_end, 128) of the tile are beyond the valid rows
#pragma unroll 1
        for (int c = c_pad; c < c_end; c += Wp)
          *reinterpret_cast<T*>(prow + (c >> 6) * kTcBlk + ((((c & 63) >> 3) ^ sw) << 4) + (c & 7) * 2) = T16<T>::from_f(0.f);
        if (c_end < kTcTile) {
          int c = max(c_end, 0);
#pragma unroll 1
          for (; c < kTcTile && (c & 7); ++c)
            *reinterpret_cast<T*>(prow + (c >> 6) * kTcBlk + ((((c & 63) >> 3) ^ sw) << 4) + (c & 7) * 2) = T16<T>::from_f(0.f);
#pragma unroll 1
          for (; c < kTcTile; c += 8)
            *reinterpret_cast<uint4*>(prow + (c >> 6) * kTcBlk + ((((c & 63) >> 3) ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
        }
      }
      // redo (warp-uniform decision) if some row of this warp saw a score more than 8 above its reference maximum
      PT_ADD(x_main, t_main);
      if (first && __any_sync(0xffffffff, mx2 > m_use + 8.f)) {
        first = false;
#ifdef BUCTD_ROLE_TIMERS
        x_redo += 1;
#endif
        continue;
      }
      break;
    }
    l_run += ls;
    if (ones && g.plant) {                                  // plant the ones column in this tile's last V block (one key row per thread)
      const int nv = (2 * j + 2) * NKC - 1, sv = nv % NS;
      { PT_BEGIN(t0); mbar_wait(&kv_full[sv], (nv / NS) & 1); PT_ADD(x_v, t0); }
      const int cl = p.c & 63;
      *reinterpret_cast<T*>(s_kv + sv * kTcBlk + etid * 128 + ((((cl >> 3)) ^ (etid & 7)) << 4) + (cl & 7) * 2) = T16<T>::from_f(1.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_full);
  }
#ifdef BUCTD_ROLE_TIMERS
  PT_ADD(x_tot, x_start);
  if (blockIdx.x == 3 && blockIdx.y == 1 && g.plant && etid == 5) {
    g_pa_cycles[4] = x_tot; g_pa_cycles[5] = x_bar; g_pa_cycles[6] = x_s; g_pa_cycles[7] = x_o; g_pa_cycles[8] = x_main;
    g_pa_cycles[9] = x_v; g_pa_cycles[10] = x_redo; g_pa_cycles[11] = ntiles; g_pa_cycles[12] = x_ld; g_pa_cycles[13] = x_ldi; g_pa_cycles[14] = x_ch;
  }
#endif
  // ---- finalise ---------------------------------------------------------------------------------------------
  mbar_wait(o_done, (ntiles - 1) & 1);
  tc_fence_after();
  const int r = m0 + row;
  const bool ok = r < p.nrows;
  const bool pad = ok && ((r % Wp) == p.W || (r / Wp) == p.H);
  if (ones) {                                               // l = O[row][C]
    uint32_t lv;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(lv) : "r"(tm_o + lane_off + p.c));
    tmem_ld_wait();
    l_run = __uint_as_float(lv);
  }
  const float inv = (ok && !pad) ? 1.f / (TRAIN ? l_run * (1.f - p.drop_p) : l_run) : 0.f;
  if (TRAIN && p.lse2 != nullptr && ok) p.lse2[(long long)row_base + r] = pad ? 0.f : m_run + log2f(l_run);
  T* go = reinterpret_cast<T*>(p.o) + ((long long)row_base + r) * p.o_cs;
#pragma unroll 1
  for (int c0 = 0; c0 < p.c; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tm_o + lane_off + c0, v);
    tmem_ld_wait();
    if (ok) {
      uint4 u0, u1;
      u0.x = T16<T>::pack2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
      u0.y = T16<T>::pack2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
      u0.z = T16<T>::pack2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
      u0.w = T16<T>::pack2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
      u1.x = T16<T>::pack2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
      u1.y = T16<T>::pack2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
      u1.z = T16<T>::pack2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
      u1.w = T16<T>::pack2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
      if (c0 + 8 <= p.c) *reinterpret_cast<uint4*>(go + c0) = u0;
      if (c0 + 16 <= p.c) *reinterpret_cast<uint4*>(go + c0 + 8) = u1;
    }
  }
}

// TRAIN: dropout of the probabilities + the log-sum-exp output (the inference instantiation carries neither: registers)
template <typename T, bool TRAIN>
__global__ void __launch_bounds__(kTcThreads, 2)
pos_att_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                  const __grid_constant__ CUtensorMap map_v, const PosAttTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // (no integer round trip: keeps the shared address space)
  const int NKC = p.nkc, NS = p.ns;
  uint8_t* s_q = smem;                       // [nkc blocks][128][128 B]
  uint8_t* s_kv = s_q + NKC * kTcBlk;        // ring of ns blocks holding the stream K(0) V(0) K(1) V(1) ... (load n -> slot n % ns)
  uint8_t* s_p = s_kv + NS * kTcBlk;         // [2 blocks: keys 0..63 | 64..127][128 queries][128 B]
  float* s_bias = reinterpret_cast<float*>(s_p + 2 * kTcBlk);   // [2][128] additive key mask (double buffered)
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_bias + 2 * kTcTile);   // [2][4] bit c%32 of word c/32: key c of the tile is masked
  uint64_t* q_full = reinterpret_cast<uint64_t*>(s_mask + 8);
  uint64_t* kv_full = q_full + 1;            // [4]
  uint64_t* kv_empty = kv_full + 4;          // [4]
  uint64_t* s_full = kv_empty + 4;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_done = p_full + 1;
  uint64_t* o_half = o_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_half + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int m0 = blockIdx.x * kTcTile;
  const int ntiles = (p.nrows + kTcTile - 1) / kTcTile;
  const int row_base = b * p.nrows;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_done, 1); mbar_init(o_half, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base, tm_o = tmem_base + kTcTile;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, (uint32_t)(NKC * kTcBlk));
      for (int kc = 0; kc < NKC; ++kc) tma_load_2d(s_q + kc * kTcBlk, &map_q, q_full, kc * 64, row_base + m0);
      uint32_t slot = 0, ph = 1;                              // waits on "empty" (starts free)
      for (int j = 0; j < ntiles; ++j) {
        for (int i = 0; i < 2 * NKC; ++i) {                   // K blocks, then V blocks of key tile j
          mbar_wait(&kv_empty[slot], ph);
          mbar_expect_tx(&kv_full[slot], kTcBlk);
          const bool is_v = i >= NKC;
          tma_load_2d(s_kv + slot * kTcBlk, is_v ? &map_v : &map_k, &kv_full[slot], (is_v ? i - NKC : i) * 64, row_base + j * kTcTile);
          if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_f16(T16<T>::kFmt, 128, kTcTile);                 // S = Q K^T, both K-major
      const uint32_t idesc_o64 = umma_idesc_f16(T16<T>::kFmt, 128, 64) | (1u << 16);       // O += P V, B (=V) MN-major
      const uint32_t idesc_ol = umma_idesc_f16(T16<T>::kFmt, 128, p.n_last) | (1u << 16);  // last V block
      const uint32_t q_addr = smem_u32(s_q);
      const uint64_t pd = umma_desc_sw128(smem_u32(s_p));
      mbar_wait(q_full, 0);
      const uint32_t kv_addr = smem_u32(s_kv);
      PT_DECL(m_k); PT_DECL(m_p); PT_DECL(m_v); PT_DECL(m_tot);
      PT_BEGIN(m_start);
      uint32_t slot = 0, ph = 0;
      for (int j = 0; j < ntiles; ++j) {
        for (int kc = 0; kc < NKC; ++kc) {                    // S = sum over channel blocks of Q_kc K_kc^T
          { PT_BEGIN(t0); mbar_wait(&kv_full[slot], ph); PT_ADD(m_k, t0); }
          tc_fence_after();
          const uint64_t qd = umma_desc_sw128(q_addr + kc * kTcBlk);
          const uint64_t kd = umma_desc_sw128(kv_addr + slot * kTcBlk);
          const int ks = (min(64, p.c - kc * 64) + 15) >> 4;
          for (int k = 0; k < ks; ++k) tc_mma_f16(tm_s, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, (kc | k) != 0);
          tc_commit(&kv_empty[slot]);
          if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
        }
        tc_commit(s_full);
        { PT_BEGIN(t0); mbar_wait(p_full, j & 1); PT_ADD(m_p, t0); }
        for (int kc = 0; kc < NKC; ++kc) {                    // O[:, 64 kc ..] += P V_kc
          { PT_BEGIN(t0); mbar_wait(&kv_full[slot], ph); PT_ADD(m_v, t0); }
          tc_fence_after();
          const uint64_t vd = umma_desc_sw128_mn(kv_addr + slot * kTcBlk);
          const uint32_t idesc_o = kc == NKC - 1 ? idesc_ol : idesc_o64;
#pragma unroll
          for (int k = 0; k < kTcTile / 16; ++k) {
            // A: P block k/4 (64 keys each), 32 B per k-step inside the 128-byte row; B: 16 key rows = 2048 B further
            const uint64_t pa = pd + (uint64_t)((k >> 2) * (kTcBlk >> 4) + (k & 3) * 2);
            tc_mma_f16(tm_o + kc * 64, pa, vd + (uint64_t)(k * (2048 >> 4)), idesc_o, (j | k) != 0);
            if (k == 3 && kc == NKC - 1) tc_commit(o_half);
          }
          tc_commit(&kv_empty[slot]);
          if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
        }
        tc_commit(o_done);
      }
#ifdef BUCTD_ROLE_TIMERS
      PT_ADD(m_tot, m_start);
      if (blockIdx.x == 3 && blockIdx.y == 1) { g_pa_cycles[0] = m_tot; g_pa_cycles[1] = m_k; g_pa_cycles[2] = m_p; g_pa_cycles[3] = m_v; }
#endif
    }
    __syncwarp();
  } else {
    PaGroup g;
    g.s_p = s_p; g.s_bias = s_bias; g.s_mask = s_mask; g.s_full = s_full; g.p_full = p_full; g.o_done = o_done; g.o_half = o_half;
    g.kv_full = kv_full; g.s_kv = s_kv; g.tm_s = tm_s; g.tm_o = tm_o; g.m0 = m0; g.etid = threadIdx.x - 64; g.bar_id = 1; g.plant = true;
    pa_softmax_role<T, TRAIN>(p, g, b, row_base, ntiles, NKC, NS);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// -------------------------------------------------------------------------------------------------------------
// Two query tiles per CTA (C <= 64: one channel block), one CTA per SM.  Each tile has its own S / O accumulators, P
// tile and group of four soft-max warps; the MMA thread alternates between them:
//     S_a(0) S_b(0) | P_a(j) -> PV_a(j), S_a(j+1) | P_b(j) -> PV_b(j), S_b(j+1) | ...
// so while group A waits for its PV + next S (11 MMAs), group B is in its exponentials and vice versa: the MUFU
// pipes of the four sub-partitions (one warp of each group on each) never wait for the tensor core.  K / V tiles
// are fetched once for both query tiles.  TMEM: S_a | S_b | O_a | O_b = 128 + 128 + 64 + 64 columns.
// -------------------------------------------------------------------------------------------------------------
constexpr int kTc2Threads = 320;

template <typename T, bool TRAIN>
__global__ void __launch_bounds__(kTc2Threads, 1)
pos_att_tc2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const PosAttTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // (no integer round trip: keeps the shared address space)
  constexpr int NS = 4;
  uint8_t* s_q = smem;                       // [2 query tiles][128][128 B]
  uint8_t* s_kv = s_q + 2 * kTcBlk;          // ring: K(0) V(0) K(1) V(1) ...
  uint8_t* s_p = s_kv + NS * kTcBlk;         // [2 groups][2 blocks][128][128 B]
  float* s_bias = reinterpret_cast<float*>(s_p + 4 * kTcBlk);             // [2 groups][2][128]
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_bias + 4 * kTcTile);   // [2 groups][2][4]
  uint64_t* q_full = reinterpret_cast<uint64_t*>(s_mask + 16);
  uint64_t* kv_full = q_full + 1;            // [4]
  uint64_t* kv_empty = kv_full + 4;          // [4]
  uint64_t* s_full = kv_empty + 4;           // [2]
  uint64_t* p_full = s_full + 2;             // [2]
  uint64_t* o_done = p_full + 2;             // [2]
  uint64_t* o_half = o_done + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_half + 2);

  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int m0 = blockIdx.x * 2 * kTcTile;
  const int ntiles = (p.nrows + kTcTile - 1) / kTcTile;
  const int row_base = b * p.nrows;

  if (warp == 0 && (threadIdx.x & 31) == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_done[i], 1); mbar_init(&o_half[i], 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, (uint32_t)(2 * kTcBlk));
      tma_load_2d(s_q, &map_q, q_full, 0, row_base + m0);
      tma_load_2d(s_q + kTcBlk, &map_q, q_full, 0, row_base + m0 + kTcTile);
      uint32_t slot = 0, ph = 1;
      for (int j = 0; j < ntiles; ++j) {
        for (int i = 0; i < 2; ++i) {                         // K, then V of key tile j
          mbar_wait(&kv_empty[slot], ph);
          mbar_expect_tx(&kv_full[slot], kTcBlk);
          tma_load_2d(s_kv + slot * kTcBlk, i ? &map_v : &map_k, &kv_full[slot], 0, row_base + j * kTcTile);
          if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_f16(T16<T>::kFmt, 128, kTcTile);
      const uint32_t idesc_o = umma_idesc_f16(T16<T>::kFmt, 128, p.n_last) | (1u << 16);
      const uint32_t q_addr = smem_u32(s_q), kv_addr = smem_u32(s_kv), p_addr = smem_u32(s_p);
      const int ks = (p.c + 15) >> 4;
      mbar_wait(q_full, 0);
      // load n of the ring stream sits in slot n % NS and completes phase (n / NS) & 1:  K(j) = load 2j, V(j) = load 2j + 1
      auto issue_s = [&](int g, int j) {
        const int n = 2 * j, slot = n % NS;
        if (g == 0) mbar_wait(&kv_full[slot], (uint32_t)(n / NS) & 1u);
        tc_fence_after();
        const uint64_t qd = umma_desc_sw128(q_addr + g * kTcBlk);
        const uint64_t kd = umma_desc_sw128(kv_addr + slot * kTcBlk);
        for (int k = 0; k < ks; ++k) tc_mma_f16(tmem_base + g * kTcTile, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k != 0);
        if (g == 1) tc_commit(&kv_empty[slot]);
        tc_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int j) {
        const int n = 2 * j + 1, slot = n % NS;
        mbar_wait(&p_full[g], (uint32_t)j & 1u);
        if (g == 0) mbar_wait(&kv_full[slot], (uint32_t)(n / NS) & 1u);
        tc_fence_after();
        const uint64_t pd = umma_desc_sw128(p_addr + g * 2 * kTcBlk);
        const uint64_t vd = umma_desc_sw128_mn(kv_addr + slot * kTcBlk);
#pragma unroll
        for (int k = 0; k < kTcTile / 16; ++k) {
          const uint64_t pa = pd + (uint64_t)((k >> 2) * (kTcBlk >> 4) + (k & 3) * 2);
          tc_mma_f16(tmem_base + 2 * kTcTile + g * 64, pa, vd + (uint64_t)(k * (2048 >> 4)), idesc_o, (j | k) != 0);
          if (k == 3) tc_commit(&o_half[g]);
        }
        if (g == 1) tc_commit(&kv_empty[slot]);
        tc_commit(&o_done[g]);
      };
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < ntiles; ++j) {
        if (p.s_first) {
          // the next S of a group goes out BEFORE its PV (P lives in shared memory, so S(j+1) may overwrite S(j) as soon
          // as P(j) is complete): the group's soft-max warps get their next scores 8 MMAs earlier
          mbar_wait(&p_full[0], (uint32_t)j & 1u);
          if (j + 1 < ntiles) issue_s(0, j + 1);
          issue_pv(0, j);
          mbar_wait(&p_full[1], (uint32_t)j & 1u);
          if (j + 1 < ntiles) issue_s(1, j + 1);
          issue_pv(1, j);
        } else {
          issue_pv(0, j);
          if (j + 1 < ntiles) issue_s(0, j + 1);
          issue_pv(1, j);
          if (j + 1 < ntiles) issue_s(1, j + 1);
        }
      }
    }
    __syncwarp();
  } else {
    const int grp = (warp - 2) >> 2;
    PaGroup g;
    g.s_p = s_p + grp * 2 * kTcBlk; g.s_bias = s_bias + grp * 2 * kTcTile; g.s_mask = s_mask + grp * 8;
    g.s_full = &s_full[grp]; g.p_full = &p_full[grp]; g.o_done = &o_done[grp]; g.o_half = &o_half[grp];
    g.kv_full = kv_full; g.s_kv = s_kv;
    g.tm_s = tmem_base + grp * kTcTile; g.tm_o = tmem_base + 2 * kTcTile + grp * 64;
    g.m0 = m0 + grp * kTcTile; g.etid = threadIdx.x - 64 - grp * 128; g.bar_id = 1 + grp; g.plant = grp == 0;
    pa_softmax_role<T, TRAIN>(p, g, b, row_base, ntiles, 1, NS);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

#ifdef BUCTD_ROLE_TIMERS
extern "C" __attribute__((visibility("default"))) int buctd_debug_pa_cycles(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_pa_cycles, sizeof(unsigned long long) * 16);
}
#endif

static size_t tc2_smem_bytes() { return 1024 + (size_t)(2 + 4 + 4) * kTcBlk + 4 * kTcTile * 4 + 512; }
static size_t tc_smem_bytes(int nkc, int ns) { return 1024 + (size_t)(nkc + ns + 2) * kTcBlk + 2 * kTcTile * 4 + 256; }

int pos_att_tc_init_device() {
  BUCTD_CUDA(cudaFuncSetAttribute(pos_att_tc_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(6, 4)));
  BUCTD_CUDA(cudaFuncSetAttribute(pos_att_tc_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(6, 4)));
  BUCTD_CUDA(cudaFuncSetAttribute(pos_att_tc_kernel<__half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(6, 4)));
  BUCTD_CUDA(cudaFuncSetAttribute(pos_att_tc_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(6, 4)));
  BUCTD_CUDA(cudaFuncSetAttribute(pos_att_tc2_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2_smem_bytes()));
  BUCTD_CUDA(cudaFuncSetAttribute(pos_att_tc2_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2_smem_bytes()));
  BUCTD_CUDA(cudaFuncSetAttribute(pos_att_tc2_kernel<__half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2_smem_bytes()));
  BUCTD_CUDA(cudaFuncSetAttribute(pos_att_tc2_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2_smem_bytes()));
  return BUCTD_OK;
}

bool pos_att_tc_supported(const buctd_pos_att_args& a) { return a.c % 16 == 0 && a.c >= 16 && a.c <= 384; }

int launch_pos_att(const buctd_pos_att_args& a, cudaStream_t st) {
  BUCTD_CHECK(a.q && a.k && a.v && a.o, BUCTD_ERR_INVALID, "pos_att: null pointer");
  BUCTD_CHECK(a.q_cs % 8 == 0 && a.k_cs % 8 == 0 && a.v_cs % 8 == 0 && a.o_cs % 8 == 0, BUCTD_ERR_INVALID,
              "pos_att: channel strides must be multiples of 8");
  BUCTD_CHECK((((uintptr_t)a.q | (uintptr_t)a.k | (uintptr_t)a.v | (uintptr_t)a.o) & 15) == 0, BUCTD_ERR_INVALID,
              "pos_att: q/k/v/o must be 16-byte aligned");
  BUCTD_CHECK(pos_att_tc_supported(a), BUCTD_ERR_INVALID, "pos_att: head dim C=%d must be a multiple of 16 in [16, 384]", a.c);
  if (a.batch == 0) return BUCTD_OK;
  return launch_pos_att_tc(a, st);
}

int launch_pos_att_tc(const buctd_pos_att_args& a, cudaStream_t st) {
  EncodeTiledFn enc = get_encode_tiled();
  BUCTD_CHECK(enc != nullptr, BUCTD_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not available");
  const int nrows = (a.h + 1) * (a.w + 1);
  const long long total = (long long)a.batch * nrows;
  const CUtensorMapDataType dt = a.dtype == BUCTD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap mq, mk, mv;
  const void* ptrs[3] = {a.q, a.k, a.v};
  const int cs[3] = {a.q_cs, a.k_cs, a.v_cs};
  CUtensorMap* maps[3] = {&mq, &mk, &mv};
  for (int i = 0; i < 3; ++i) {
    cuuint64_t dims[2] = {(cuuint64_t)a.c, (cuuint64_t)total};
    cuuint64_t strides[1] = {(cuuint64_t)cs[i] * 2};
    cuuint32_t box[2] = {64, kTcTile};
    cuuint32_t es[2] = {1, 1};
    CUresult cr = enc(maps[i], dt, 2, const_cast<void*>(ptrs[i]), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BUCTD_CHECK(cr == CUDA_SUCCESS, BUCTD_ERR_CUDA, "cuTensorMapEncodeTiled(pos_att %d) failed: %d", i, (int)cr);
  }
  PosAttTcParams p;
  p.nrows = nrows; p.W = a.w; p.H = a.h; p.c = a.c; p.o_cs = a.o_cs;
  p.nkc = (a.c + 63) / 64;
  // C <= 64 (one channel block) and at least two query tiles: two query tiles per CTA (pos_att_tc2_kernel, 4 ring slots)
  static const int two_env = getenv("BUCTD_PA_TWO") ? atoi(getenv("BUCTD_PA_TWO")) : 1;     // 0: always one query tile per CTA
  const bool two = p.nkc == 1 && nrows > kTcTile && two_env != 0;
  p.ns = (p.nkc == 1 && !two) ? 3 : 4;
  // the ones column needs a free channel slot in the last 64-channel block, and the last V block of a tile must be able
  // to land while the soft-max warps still hold P back (ring >= nkc + 2 slots): C = 16 .. 48, 80 .. 112
  p.ones = (a.c % 64 != 0 && p.nkc + 2 <= p.ns && !(a.drop_p > 0.f)) ? 1 : 0;     // (dropout: the ones column would sum the dropped P)
  BUCTD_CHECK(a.drop_p >= 0.f && a.drop_p < 1.f, BUCTD_ERR_INVALID, "pos_att: drop_p %f out of [0, 1)", a.drop_p);
  p.drop_p = a.drop_p; p.drop_seed = a.drop_seed; p.lse2 = a.lse2; p.seed_xor = a.seed_xor;
  const int c_last = a.c - 64 * (p.nkc - 1);
  p.n_last = p.ones ? (c_last + 1 + 15) / 16 * 16 : c_last;
  p.n2 = 64 * (p.nkc - 1) + p.n_last;
  p.tmem_cols = kTcTile + p.n2 <= 256 ? 256 : 512;
  p.sc = 1.4426950408889634f / sqrtf((float)a.c);
  p.o = a.o;
  cudaError_t e;
  const bool train = a.drop_p > 0.f || a.lse2 != nullptr;
  const int ntiles = (nrows + kTcTile - 1) / kTcTile;
  static const int s_first_env = getenv("BUCTD_PA_SFIRST") ? atoi(getenv("BUCTD_PA_SFIRST")) : 1;
  p.s_first = s_first_env;
  p.dbg = getenv("BUCTD_PA_DBG") ? atoi(getenv("BUCTD_PA_DBG")) : 0;
  if (two) {
    dim3 grid((ntiles + 1) / 2, a.batch);
    const size_t smem = tc2_smem_bytes();
    if (a.dtype == BUCTD_F16) e = train ? launch_pdl(pos_att_tc2_kernel<__half, true>, grid, dim3(kTc2Threads), smem, st, mq, mk, mv, p)
                                        : launch_pdl(pos_att_tc2_kernel<__half, false>, grid, dim3(kTc2Threads), smem, st, mq, mk, mv, p);
    else e = train ? launch_pdl(pos_att_tc2_kernel<__nv_bfloat16, true>, grid, dim3(kTc2Threads), smem, st, mq, mk, mv, p)
                   : launch_pdl(pos_att_tc2_kernel<__nv_bfloat16, false>, grid, dim3(kTc2Threads), smem, st, mq, mk, mv, p);
  } else {
    dim3 grid(ntiles, a.batch);
    const size_t smem = tc_smem_bytes(p.nkc, p.ns);
    if (a.dtype == BUCTD_F16) e = train ? launch_pdl(pos_att_tc_kernel<__half, true>, grid, dim3(kTcThreads), smem, st, mq, mk, mv, p)
                                        : launch_pdl(pos_att_tc_kernel<__half, false>, grid, dim3(kTcThreads), smem, st, mq, mk, mv, p);
    else e = train ? launch_pdl(pos_att_tc_kernel<__nv_bfloat16, true>, grid, dim3(kTcThreads), smem, st, mq, mk, mv, p)
                   : launch_pdl(pos_att_tc_kernel<__nv_bfloat16, false>, grid, dim3(kTcThreads), smem, st, mq, mk, mv, p);
  }
  BUCTD_CHECK(e == cudaSuccess, BUCTD_ERR_CUDA, "pos_att_tc launch failed: %s", cudaGetErrorString(e));
  BUCTD_CUDA(cudaGetLastError());
  return BUCTD_OK;
}

}  // namespace buctd
