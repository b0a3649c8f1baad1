// cmg_device.cuh -- sm_100a kernels of the Ising SGC Metropolis path.
//
// Data layout in HBM: every lattice ("chain") is held as
// two int8 checkerboard colour planes.  A site (i,j,k) has colour
// c = (i+j+k)&1 and lives in plane c at plane index
//     q = (i>>1) + h*(j + n1*k),  h = n0/2
// so i = 2*(q%h) + ((j+k+c)&1).  A plane byte is the occupation index
// b = (1+s)/2 in {0,1} of the reference's +1/-1 occupation
// (include/casm/monte/ising_cpp/model.hh:48).  With this layout the 2*dim
// neighbours of a colour-c site all sit in plane 1-c, at the SAME p for the
// j/k neighbours and at p and p+-1 for the two i neighbours.
//
// The physics is table-driven: for each chain the host builds dE and
// exp(-dE*beta) for the 2*(2*dim+1) possible (b, n_up) cases with the
// reference's exact double-precision expression order; kernels index the table
// with idx = 2*n_up + b.  No floating point is evaluated in the hot kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

// build-time tuning knobs (tools/ab_build.py compiles variants side by side)
#ifndef CMG_BULK_CTAS
#define CMG_BULK_CTAS 4  // resident CTAs per SM the bulk kernels are compiled for
#endif
#ifndef CMG_BULK_PAIR
#define CMG_BULK_PAIR 1  // bulk kernels: 1 = pair table, 0 = two lane tables
#endif

namespace cmg {

constexpr int kMaxIdx = 14;  // 2*(2*3+1)

struct ChainTables {
  double dE[16];
  double prob[16];
  uint32_t thr_m1[16];  // checkerboard: accept iff philox_u32 <= thr_m1[idx]
  double J, mu, temperature, beta;
  int valid;
  // bit idx set: exp(-dE*beta) underflowed to 0, the reference's `rand < prob`
  // (methods/metropolis.hh:33) can never accept; thr_m1[idx] is 0 for such an
  // entry, so the fast compare can only tie and the exact compare rejects
  uint32_t never_mask;
  int pad[2];
};

struct LatticeView {
  uint8_t *planes;       // chain 0, colour 0
  long long plane_stride;  // bytes between the two colour planes of a chain
  long long chain_stride;  // bytes between chains
  int h, n1, n2, dim;      // h = n0/2 (planes layout)
  // slab decomposition (2-d): global column index of local column 0 and the
  // halo columns standing in for local columns -1 and n1.  halo_lo/hi[c] point
  // at h bytes of plane-c data; null => periodic wrap inside this lattice.
  long long col_offset;
  const uint8_t *halo_lo[2];
  const uint8_t *halo_hi[2];
  // fused halo push: where to store this slab's freshly updated boundary
  // columns (the neighbour's halo buffers, possibly peer memory); null => none
  uint8_t *push_lo[2];  // our column 0      -> low neighbour's halo_hi
  uint8_t *push_hi[2];  // our column n1-1   -> high neighbour's halo_lo
  // cross-GPU ordering of the fused push (all null when not used):
  // wait_flag[side] lives in OUR memory and is raised by the neighbour on that
  // side when its half-sweep has finished pushing; signal_flag[side] is the
  // neighbour's wait flag (peer memory).  Values count finished half-sweeps.
  const unsigned long long *wait_flag[2];
  unsigned long long *signal_flag[2];
  unsigned int *done_counter;   // [3]: CTAs of this launch that have finished; edge CTAs done on side 0 / 1
  int edge_mode;                // slab: the boundary columns are swept, pushed and signalled first (below)
  unsigned long long epoch;     // number of fused half-sweeps stepped before this one
  unsigned int *error;          // sticky error word of the context (bit 2: a neighbour wait timed out)
};

struct SweepArgs {
  LatticeView L;
  const ChainTables *tabs;
  unsigned long long *n_accept;  // [chain]
  long long *sb;                 // sample slot of chain 0: {S,B}; null if !SAMPLE
  long long sb_chain_stride;     // in long long units
  unsigned long long pass;
  // Philox round keys k_r = seed + r*(W0,W1), r = 0..9: launch-uniform, so they
  // are read straight from the kernel-parameter constant bank
  uint32_t rk[20];
  int colour;
  int js;  // columns per thread strip (bulk kernels)
  // bulk3d, when > 0: n_strips balanced strips per layer (even starts, lengths
  // within two columns of each other) instead of uniform js-column strips, and
  // the lane groups of a warp take layers k, k+2, ... of ONE strip (pair_layers)
  int n_strips;
  int pair_layers;
  int chain_offset;  // global index of chain 0 (chains sharded over several contexts)
  // Chained half-sweeps (bulk2d / bulk3d, one context, consecutive launches of one run):
  // every CTA publishes hs_epoch in hs_flags[chain][cta] when its columns are written;
  // with hs_wait a thread waits for the CTAs of the previous half-sweep that wrote what
  // it reads (and read what it writes) instead of for the whole previous grid.
  unsigned int *hs_flags;
  uint32_t hs_epoch;
  int hs_wait;
  unsigned int *error;  // sticky error word (bounded waits)
};

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Counter-based: the
// stream is a pure function of (site group, pass, colour, seed, chain), so the
// trajectory does not depend on launch geometry or on the number of GPUs.
// ---------------------------------------------------------------------------
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

// rk = the ten round-key pairs (key + r*(W0,W1)).  R = 10 is the published default;
// R = 7 (cmg_set_philox_rounds) is the fewest rounds of Philox4x32 that pass BigCrush
// (Salmon et al., table 2) -- an opt-in, never chosen by the library itself.
template <int R = 10>
__device__ __forceinline__ uint4 philox4x32(uint4 c, const uint32_t *rk) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    unsigned long long p0 = (unsigned long long)kPhiloxM0 * c.x;
    unsigned long long p1 = (unsigned long long)kPhiloxM1 * c.z;
    uint4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ rk[2 * r];
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ rk[2 * r + 1];
    n.w = (uint32_t)p0;
    c = n;
  }
  return c;
}

// Random words of site group `group` (8 consecutive plane indices, one 16-bit
// lane each) of `chain` in half-sweep (pass, colour): counter = {lo32(group),
// (hi32(group)&0xff) | chain<<8, lo32(pass), hi32(pass)<<2 | refine<<1 | colour},
// key = seed.  refine = 0: the lane r16 that supplies the leading 16 bits of each
// site's uniform (rotated, see accept_mask4_fast); refine = 1: the trailing 16
// bits r16' (only evaluated on a tie).
template <int R = 10>
__device__ __forceinline__ uint4 site_group_random(unsigned long long group,
                                                   uint32_t chain_word,
                                                   unsigned long long pass,
                                                   int colour, int refine,
                                                   const uint32_t *rk) {
  uint4 c;
  c.x = (uint32_t)group;
  c.y = ((uint32_t)(group >> 32) & 0xffu) | chain_word;
  c.z = (uint32_t)pass;
  c.w = ((uint32_t)(pass >> 32) << 2) | ((uint32_t)refine << 1) | (uint32_t)colour;
  return philox4x32<R>(c, rk);
}

// All sweep kernels use one dynamic shared-memory buffer, declared at
// namespace scope so that every access is a plain LDS (no generic-address
// conversion):
//   [0, 64)                 the chain's acceptance table thr_m1, 16 x u32 (exact path)
//   [kSmemLaneLo, +64)      fast-compare lane G[idx] of every entry, in the low half
//   [kSmemLaneHi, +64)      the same lane in the high half (G[idx] << 16)
//   [kSmemNever, +4)        never_mask of the chain (exact path only)
//   [kSmemPair, +14*1024)   pair table: entry (A, B) = G[A] | G[B] << 16 at byte
//                           offset 4*A + 1024*B, so the 16-bit value formed by two
//                           neighbouring index bytes (each holding 4*index) IS the
//                           byte offset of the pair: one LDS per two sites
//   [kSmemTile, ...)        tile data (k_tile2d) / staging ring (bulk kernels)
extern __shared__ __align__(16) unsigned char cmg_smem[];
constexpr int kSmemLaneLo = 64;
constexpr int kSmemLaneHi = 128;
constexpr int kSmemNever = 192;  // u32: the chain's never_mask
constexpr int kSmemSmall = 256;  // kernels that only use the 16-entry tables
constexpr int kSmemPair = 1024;
constexpr int kPairCopy = 64;  // byte offset of the odd lanes' copy of every pair-table row
constexpr int kSmemTile = kSmemPair + 14 * 1024;

// Fast-compare lane of a table entry (see accept_mask4_fast): 0x7FFF + the
// leading 15 bits of the threshold, 0xFFFF for an entry that always accepts.
__device__ __forceinline__ uint32_t fast_lane(uint32_t thr_m1) {
  return thr_m1 == 0xFFFFFFFFu ? 0xFFFFu : 0x7FFFu + (thr_m1 >> 17);
}

__device__ __forceinline__ void load_accept_table(const ChainTables *tab, bool with_pairs) {
  if (threadIdx.x < 16) {
    const uint32_t thr = tab->thr_m1[threadIdx.x];
    reinterpret_cast<uint32_t *>(cmg_smem)[threadIdx.x] = thr;
    reinterpret_cast<uint32_t *>(cmg_smem + kSmemLaneLo)[threadIdx.x] = fast_lane(thr);
    reinterpret_cast<uint32_t *>(cmg_smem + kSmemLaneHi)[threadIdx.x] = fast_lane(thr) << 16;
    if (threadIdx.x == 0) *reinterpret_cast<uint32_t *>(cmg_smem + kSmemNever) = tab->never_mask;
  }
  if (!with_pairs) return;
  for (int e = threadIdx.x; e < 14 * 14; e += blockDim.x) {
    const int A = e % 14, B = e / 14;
    const uint32_t pair = fast_lane(tab->thr_m1[A]) | (fast_lane(tab->thr_m1[B]) << 16);
    *reinterpret_cast<uint32_t *>(cmg_smem + kSmemPair + 4 * A + 1024 * B) = pair;
    // second copy 16 banks further on (rows only use 56 of their 1024 bytes): odd
    // lanes look up there, which halves the bank conflicts of the random lookups
    *reinterpret_cast<uint32_t *>(cmg_smem + kSmemPair + kPairCopy + 4 * A + 1024 * B) = pair;
  }
}
// byte_off = 4 * table index
__device__ __forceinline__ uint32_t thr_at(uint32_t byte_off) {
  return *reinterpret_cast<const uint32_t *>(cmg_smem + byte_off);
}
__device__ __forceinline__ bool never_at(uint32_t byte_off) {
  return ((*reinterpret_cast<const uint32_t *>(cmg_smem + kSmemNever) >> (byte_off >> 2)) & 1u) != 0u;
}
__device__ __forceinline__ uint32_t lane_lo_at(uint32_t byte_off) {
  return *reinterpret_cast<const uint32_t *>(cmg_smem + kSmemLaneLo + byte_off);
}
__device__ __forceinline__ uint32_t lane_hi_at(uint32_t byte_off) {
  return *reinterpret_cast<const uint32_t *>(cmg_smem + kSmemLaneHi + byte_off);
}
// pair_off = 4*indexA + 1024*indexB
__device__ __forceinline__ uint32_t lane_pair_at(uint32_t pair_off) {
  return *reinterpret_cast<const uint32_t *>(cmg_smem + kSmemPair + pair_off);
}

// Acceptance of 4 sites packed in a word.  Each site's uniform is the 32-bit
// integer R = rotl16(r16, 1) << 16 | r16' (r16, r16': the site's 16-bit Philox
// lanes of the leading / refinement call) and the site is flipped iff
// R <= thr (thr = thr_m1 of its table entry).  The leading 15 bits of R are
// a = r16 & 0x7FFF, so with T = thr >> 17:
//   a < T   -> accepted whatever the remaining 17 bits are,
//   a > T   -> rejected,
//   a == T  -> a tie (probability 2^-15 per site) that needs the exact compare.
// Two sites are compared per 32-bit subtraction: their lanes G = 0x7FFF + T sit
// in the two halves of a word and x = G - a never borrows across the halves;
// x >= 0x8000 (sign bit of the half) iff accepted, x == 0x7FFF iff tie.  Entries
// that always accept have G = 0xFFFF and can never tie.
// idx4e: 4 * table index of each site in its byte; r01 / r23: the Philox words
// holding the r16 of sites (0,1) / (2,3) in their (low, high) halves.
// tmax: running per-half signed maximum (start at 0): a half equals 0x7FFF iff
// some site tied.  Returns 0xFF in the byte of every accepted site (ties count
// as rejected here; the caller redoes the vector exactly when one occurred).
// PAIR: lanes come two at a time from the pair table; otherwise one LDS per
// site from the two 16-entry lane tables (bank-conflict free) merged by the add.
template <bool PAIR>
__device__ __forceinline__ uint32_t accept_mask4_fast(uint32_t idx4e, uint32_t r01,
                                                      uint32_t r23, uint32_t &tmax) {
  const uint32_t a01 = r01 & 0x7fff7fffu, a23 = r23 & 0x7fff7fffu;
  uint32_t x01, x23;
  if (PAIR) {
    x01 = lane_pair_at(idx4e & 0xffffu) - a01;
    x23 = lane_pair_at(idx4e >> 16) - a23;
  } else {
    x01 = lane_lo_at(idx4e & 0xffu) + lane_hi_at(__byte_perm(idx4e, 0u, 0x4441u)) - a01;
    x23 = lane_lo_at(__byte_perm(idx4e, 0u, 0x4442u)) + lane_hi_at(idx4e >> 24) - a23;
  }
  tmax = __vimax3_s16x2(tmax, x01, x23);
  uint32_t m;
  // bytes 1 and 3 of x01, x23 with their sign bit replicated over the byte
  asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(m) : "r"(x01), "r"(x23));
  return m;
}
__device__ __forceinline__ bool any_tie(uint32_t tmax) {
  return ((tmax + 0x00010001u) & 0x80008000u) != 0u;
}

// Exact form with both halves (tie path and the generic kernel).
__device__ __forceinline__ bool accept_exact(uint32_t r16, uint32_t r16b, uint32_t thr) {
  const uint32_t lead = ((r16 << 1) | (r16 >> 15)) & 0xffffu;  // rotl16(r16, 1)
  return ((lead << 16) | r16b) <= thr;
}
// exact decision for the table entry at byte_off = 4 * index
__device__ __forceinline__ bool accept_exact_at(uint32_t r16, uint32_t r16b, uint32_t byte_off) {
  return accept_exact(r16, r16b, thr_at(byte_off)) && !never_at(byte_off);
}
__device__ __forceinline__ uint32_t lane16(uint4 v, int lane) {
  const uint32_t w = (lane >> 1) == 0 ? v.x : (lane >> 1) == 1 ? v.y : (lane >> 1) == 2 ? v.z : v.w;
  return (lane & 1) ? (w >> 16) : (w & 0xffffu);
}

// sum of the four bytes of a word
__device__ __forceinline__ int bytesum(uint32_t w) {
  return (int)__dp4a(w, 0x01010101u, 0u);
}

// block-wide sums -> one atomic per CTA and quantity
template <int NT>
__device__ __forceinline__ void block_accumulate(unsigned int acc, long long s_ones,
                                                 long long bsum, bool sample,
                                                 unsigned long long *n_accept,
                                                 long long *sb) {
  __shared__ unsigned int sh_acc[NT / 32];
  __shared__ long long sh_s[NT / 32];
  __shared__ long long sh_b[NT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  acc = __reduce_add_sync(0xffffffffu, acc);
  if (sample) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_ones += __shfl_xor_sync(0xffffffffu, s_ones, o);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    }
  }
  if (lane == 0) {
    sh_acc[warp] = acc;
    sh_s[warp] = s_ones;
    sh_b[warp] = bsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long a = 0;
    long long s = 0, b = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
      a += sh_acc[w];
      s += sh_s[w];
      b += sh_b[w];
    }
    if (a) atomicAdd(n_accept, a);
    if (sample) {
      atomicAdd((unsigned long long *)&sb[0], (unsigned long long)s);
      atomicAdd((unsigned long long *)&sb[1], (unsigned long long)b);
    }
  }
}

// ---------------------------------------------------------------------------
// k_halfsweep_generic: any even extents, 2-d or 3-d.  One thread per group of
// four consecutive plane indices (one Philox call).  Byte accesses; this is the
// correctness baseline the tuned kernels are cross-checked against.
//
// SAMPLE: the launch is the colour-1 half-sweep that completes a sampled pass;
// it accumulates ones(plane0)+ones(plane1) and B = sum over colour-1 sites of
// s*(sum of neighbour s) (every bond joins the two colours exactly once).
// The host turns ones into S = 2*ones - N.
// ---------------------------------------------------------------------------
template <bool SAMPLE, int R = 10>
__global__ void __launch_bounds__(128) k_halfsweep_generic(SweepArgs A) {
  const LatticeView &L = A.L;
  const int chain = blockIdx.y;
  load_accept_table(A.tabs + chain, false);
  __syncthreads();

  uint8_t *C = L.planes + (long long)chain * L.chain_stride +
               (long long)A.colour * L.plane_stride;
  const uint8_t *O = L.planes + (long long)chain * L.chain_stride +
                     (long long)(1 - A.colour) * L.plane_stride;
  const long long plane_size = (long long)L.h * L.n1 * L.n2;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int z = 2 * L.dim;

  unsigned int acc = 0;
  long long ones = 0, bsum = 0;
  if (8 * g < plane_size) {
    const uint4 ra = site_group_random<R>((unsigned long long)g, (uint32_t)(chain + A.chain_offset) << 8, A.pass,
                                          A.colour, 0, A.rk);
    const uint4 rb = site_group_random<R>((unsigned long long)g, (uint32_t)(chain + A.chain_offset) << 8, A.pass,
                                          A.colour, 1, A.rk);
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const long long q = 8 * g + w;
      if (q >= plane_size) break;
      const int p = (int)(q % L.h);
      const long long jk = q / L.h;
      const int j = (int)(jk % L.n1);
      const int k = (int)(jk / L.n1);
      const int par = (j + k + A.colour) & 1;  // i = 2p + par
      const int jm = (j == 0) ? L.n1 - 1 : j - 1, jp = (j == L.n1 - 1) ? 0 : j + 1;
      const long long rowk = (long long)L.n1 * k;
      int n_up = O[p + (long long)L.h * (jm + rowk)] + O[p + (long long)L.h * (jp + rowk)] +
                 O[p + (long long)L.h * (j + rowk)];
      // the other i-neighbour: p-1 if i even, p+1 if i odd (periodic in p)
      const int ps = par ? ((p == L.h - 1) ? 0 : p + 1) : ((p == 0) ? L.h - 1 : p - 1);
      n_up += O[ps + (long long)L.h * (j + rowk)];
      if (L.dim == 3) {
        const int km = (k == 0) ? L.n2 - 1 : k - 1, kp = (k == L.n2 - 1) ? 0 : k + 1;
        n_up += O[p + (long long)L.h * (j + (long long)L.n1 * km)] +
                O[p + (long long)L.h * (j + (long long)L.n1 * kp)];
      }
      int b = C[q];
      if (accept_exact_at(lane16(ra, w), lane16(rb, w), 4u * (2 * n_up + b))) {
        b ^= 1;
        C[q] = (uint8_t)b;
        ++acc;
      }
      if (SAMPLE) {
        ones += b + O[q];
        bsum += (2 * b - 1) * (2 * n_up - z);
      }
    }
  }
  block_accumulate<128>(acc, ones, bsum, SAMPLE, A.n_accept + chain,
                        SAMPLE ? A.sb + (long long)chain * A.sb_chain_stride : nullptr);
}

// ---------------------------------------------------------------------------
// k_halfsweep_bulk2d: the production 2-d kernel.  Requires n0 % 32 == 0 (so a
// plane column is a whole number of 16-byte vectors) and n1 even.
//
// A thread owns one 16-byte vector (16 same-colour sites, consecutive p) and
// walks a strip of `js` columns.  The three opposite-colour columns j-1, j,
// j+1 needed by column j are kept in registers as a rolling window, so moving
// to the next column costs one 16-byte load of the opposite plane, one of the
// own plane, the word holding the p+-1 neighbour across the vector edge (an L1
// hit), and one 16-byte store: 3 B of traffic per attempted flip.  The loads
// are staged four columns ahead through a per-thread cp.async ring in shared
// memory (see below).  Neighbour counts are formed four sites at a time with
// plain 32-bit adds (bytes are 0/1, sums <= 4 never carry).  Two Philox calls
// give the 16 leading lanes.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld16(const uint8_t *p) {
  return *reinterpret_cast<const uint4 *>(p);
}
__device__ __forceinline__ uint4 ld16_nc(const uint8_t *p) {
  return __ldg(reinterpret_cast<const uint4 *>(p));
}
__device__ __forceinline__ uint4 ld16_cg(const uint8_t *p) {
  return __ldcg(reinterpret_cast<const uint4 *>(p));
}

// out byte k = in byte k-1, byte 0 <- lo (a single byte value)
__device__ __forceinline__ uint4 shift_up_1(uint4 v, uint32_t lo) {
  uint4 o;
  o.x = (v.x << 8) | lo;
  o.y = __funnelshift_l(v.x, v.y, 8);
  o.z = __funnelshift_l(v.y, v.z, 8);
  o.w = __funnelshift_l(v.z, v.w, 8);
  return o;
}
// out byte k = in byte k+1, byte 15 <- hi
__device__ __forceinline__ uint4 shift_down_1(uint4 v, uint32_t hi) {
  uint4 o;
  o.x = __funnelshift_r(v.x, v.y, 8);
  o.y = __funnelshift_r(v.y, v.z, 8);
  o.z = __funnelshift_r(v.z, v.w, 8);
  o.w = (v.w >> 8) | (hi << 24);
  return o;
}

// Rarer still: the chain has entries that can never accept (exp(-dE*beta) underflowed to 0,
// thr_m1 = 0) and a vector with a tie is being redone: clear the mask bytes of such sites.
__device__ __noinline__ uint4 clear_never16(uint4 m, uint4 idx4e) {
  const uint32_t never = *reinterpret_cast<const uint32_t *>(cmg_smem + kSmemNever);
  uint32_t mw[4] = {m.x, m.y, m.z, m.w};
  const uint32_t iw[4] = {idx4e.x, idx4e.y, idx4e.z, idx4e.w};
#pragma unroll
  for (int w = 0; w < 4; ++w)
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if ((never >> (((iw[w] >> (8 * k)) & 0xffu) >> 2)) & 1u) mw[w] &= ~(0xffu << (8 * k));
  return make_uint4(mw[0], mw[1], mw[2], mw[3]);
}

// Rare path: some site of a 16-site vector tied on its leading 15 bits.  Redo
// all 16 decisions exactly with both halves (regenerating the leading words so
// the hot path does not have to keep them alive).
template <int R = 10>
__device__ __noinline__ uint4 resolve_ties16(uint4 idx4e, unsigned long long group0,
                                             unsigned long long pass, int colour,
                                             uint32_t chain_word, uint32_t key0, uint32_t key1) {
  // the round keys are rebuilt here (by value) so that the kernel parameter block
  // never has its address taken -- that would copy it to local memory at entry
  uint32_t rk[20];
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    rk[2 * r] = key0 + (uint32_t)r * kPhiloxW0;
    rk[2 * r + 1] = key1 + (uint32_t)r * kPhiloxW1;
  }
  const uint32_t iw[4] = {idx4e.x, idx4e.y, idx4e.z, idx4e.w};
  uint32_t m[4];
  for (int half = 0; half < 2; ++half) {
    const uint4 r = site_group_random<R>(group0 + half, chain_word, pass, colour, 0, rk);
    const uint4 q = site_group_random<R>(group0 + half, chain_word, pass, colour, 1, rk);
#pragma unroll
    for (int ww = 0; ww < 2; ++ww) {
      const int w = 2 * half + ww;
      uint32_t mm = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int lane = 4 * ww + k;
        // (entries that can never accept are taken out below, by a subroutine of their own:
        // testing them here, per site, changes the register allocation of the CALLERS' hot
        // loops -- measured 2.5-4 % on k_ring2d and k_halfsweep_bulk2d)
        mm |= accept_exact(lane16(r, lane), lane16(q, lane), thr_at((iw[w] >> (8 * k)) & 0xffu)) ? (0xffu << (8 * k)) : 0u;
      }
      m[w] = mm;
    }
  }
  if (*reinterpret_cast<const uint32_t *>(cmg_smem + kSmemNever) != 0u)
    return clear_never16(make_uint4(m[0], m[1], m[2], m[3]), idx4e);
  return make_uint4(m[0], m[1], m[2], m[3]);
}

// Per-thread accumulators.  Sampling terms are kept as raw byte sums and turned
// into (ones, B) once, at the end of the thread's work:
//   c1   = #(b = 1) among the updated sites,   opp = #(b = 1) of the facing sites
//   u7   = sum over updated sites of (b ? n : 7 - n),   n = # of +1 neighbours
//   B    = sum (2b-1)(2n-z) = 2*u7 - 2*(7-z)*(sites - c1) - z*sites
struct Accum {
  unsigned int acc;
  unsigned int c1, opp, u7, sites;
};
__device__ __forceinline__ void accum_finish(const Accum &a, int z, long long &ones,
                                             long long &bsum) {
  ones = (long long)a.c1 + (long long)a.opp;
  bsum = 2ll * a.u7 - 2ll * (7 - z) * ((long long)a.sites - (long long)a.c1) -
         (long long)z * a.sites;
}

// Update 16 sites (one 16-byte vector of a colour plane); returns the new
// centre vector.  group0 = plane index of the first site >> 3.
template <bool SAMPLE, bool PAIR = false, int R = 10>
__device__ __forceinline__ uint4 update16(uint4 ce, uint4 om, uint4 oc, uint4 op,
                                          uint4 side, unsigned long long group0,
                                          unsigned long long pass, int colour,
                                          uint32_t chain_word, const uint32_t *rk,
                                          Accum &a) {
  uint32_t cw[4] = {ce.x, ce.y, ce.z, ce.w};
  const uint32_t nw[4] = {om.x + oc.x + op.x + side.x, om.y + oc.y + op.y + side.y,
                          om.z + oc.z + op.z + side.z, om.w + oc.w + op.w + side.w};
  const uint32_t ow[4] = {oc.x, oc.y, oc.z, oc.w};
  const uint4 ra = site_group_random<R>(group0, chain_word, pass, colour, 0, rk);
  const uint4 rb = site_group_random<R>(group0 + 1, chain_word, pass, colour, 0, rk);
  const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
  uint32_t idx[4], m[4], tmax = 0;
  // PAIR: bytes 0 and 2 also carry the lane's pair-table copy (0 or kPairCopy)
  const uint32_t copy2 = PAIR ? (threadIdx.x & 1u) * (uint32_t)(kPairCopy | (kPairCopy << 16)) : 0u;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    idx[w] = ((nw[w] + nw[w] + cw[w]) << 2) + copy2;  // 4 * (2*n_up + b) per byte
    m[w] = accept_mask4_fast<PAIR>(idx[w], rw[2 * w], rw[2 * w + 1], tmax);
  }
  if (any_tie(tmax)) {  // a tie somewhere in these 16 sites (probability 16 * 2^-15)
    const uint4 mm = resolve_ties16<R>(make_uint4(idx[0] - copy2, idx[1] - copy2, idx[2] - copy2,
                                               idx[3] - copy2),
                                    group0, pass, colour, chain_word, rk[0], rk[1]);
    m[0] = mm.x;
    m[1] = mm.y;
    m[2] = mm.z;
    m[3] = mm.w;
  }
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    cw[w] ^= m[w] & 0x01010101u;
    a.acc = (unsigned int)__dp4a((int)m[w], (int)0xffffffffu, (int)a.acc);  // (-1)*(-1) per accepted site
  }
  if (SAMPLE) {
    // Byte sums of the four words first (plain adds: a byte holds at most
    // 4 * 7), one dot product per quantity afterwards.  Per byte
    // t = 0x80 - b is 0x80 (b = 0) or 0x7f (b = 1), so ~t & 7 is 7 where b = 0:
    // x = n ^ (~t & 7) = (b ? n : 7 - n) in one LOP3, with no multiply.
    uint32_t x[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t t = 0x80808080u - cw[w];
      x[w] = nw[w] ^ (~t & 0x07070707u);
    }
    a.u7 = __dp4a((x[0] + x[1]) + (x[2] + x[3]), 0x01010101u, a.u7);
    a.c1 = __dp4a((cw[0] + cw[1]) + (cw[2] + cw[3]), 0x01010101u, a.c1);
    a.opp = __dp4a((ow[0] + ow[1]) + (ow[2] + ow[3]), 0x01010101u, a.opp);
    a.sites += 16;
  }
  return make_uint4(cw[0], cw[1], cw[2], cw[3]);
}

// ---- cross-GPU ordering for slab decomposition ------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned int kErrRingEdge = 1u, kErrRingCopy = 2u, kErrSlabWait = 4u, kErrChain = 8u;
constexpr unsigned long long kSlabWaitNs = 20ull * 1000ull * 1000ull * 1000ull;  // 20 s
// before reading halos: both neighbours must have finished `epoch` half-sweeps.
// The wait is bounded: a neighbour that never arrives (dead rank, sequence
// mismatch) raises the context's sticky error word instead of hanging the GPU.
__device__ __forceinline__ void slab_wait_neighbours(const LatticeView &L) {
  if (L.epoch == 0 || (!L.wait_flag[0] && !L.wait_flag[1])) return;
  if (threadIdx.x == 0) {
    for (int side = 0; side < 2; ++side)
      if (L.wait_flag[side] && ld_acquire_sys(L.wait_flag[side]) < L.epoch) {
        // a wait that already timed out is not repeated by every later CTA and launch
        if (L.error && (*(volatile unsigned int *)L.error & kErrSlabWait)) break;
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_sys(L.wait_flag[side]) < L.epoch) {
          __nanosleep(64);
          if (globaltimer_ns() - t0 > kSlabWaitNs) {
            if (L.error) atomicOr(L.error, kErrSlabWait);
            break;
          }
        }
      }
  }
  __syncthreads();
}
// after the last store: the last CTA of the launch raises the neighbours' flags.
// `pushed`: this thread stored into a neighbour's halo (only those threads pay
// for a system-scope fence; the CTA barrier + the counter chain order the rest).
__device__ __forceinline__ void slab_signal_neighbours(const LatticeView &L, bool pushed) {
  if (!L.done_counter) return;
  if (pushed) __threadfence_system();  // this thread's peer stores are visible system-wide
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int prev = atomicAdd(L.done_counter, 1u);
    if (prev == total - 1) {
      *L.done_counter = 0;
      __threadfence_system();
      for (int side = 0; side < 2; ++side)
        if (L.signal_flag[side]) st_release_sys(L.signal_flag[side], L.epoch + 1);
    }
  }
}

// Edge mode (every CTA lies inside one strip, at least two strips): only the CTAs of
// the first and of the last strip touch a halo.  They come first in the grid, sweep
// the two boundary columns of their strip before the rest of it, push them and raise
// the neighbour's flag at once -- one column-time into the half-sweep instead of at
// the end of the launch -- so a neighbour that runs a whole half-sweep behind or
// ahead never waits.  flag[side] counts the half-sweeps whose edge CTAs on that side
// are done; a CTA waits for flag >= epoch before it reads that halo, which also
// orders its push (into the buffer the neighbour read one half-sweep earlier).
__device__ __forceinline__ void slab_wait_side(const LatticeView &L, int side) {
  if (L.epoch == 0 || !L.wait_flag[side]) return;
  if (threadIdx.x == 0 && ld_acquire_sys(L.wait_flag[side]) < L.epoch &&
      !(L.error && (*(volatile unsigned int *)L.error & kErrSlabWait))) {
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(L.wait_flag[side]) < L.epoch) {
      __nanosleep(64);
      if (globaltimer_ns() - t0 > kSlabWaitNs) {
        if (L.error) atomicOr(L.error, kErrSlabWait);
        break;
      }
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void slab_signal_side(const LatticeView &L, int side, unsigned int n_edge_ctas) {
  if (!L.done_counter || !L.signal_flag[side]) return;
  __threadfence_system();  // this thread's peer stores are visible system-wide
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(L.done_counter + 1 + side, 1u);
    if (prev == n_edge_ctas - 1) {
      L.done_counter[1 + side] = 0;
      __threadfence_system();
      st_release_sys(L.signal_flag[side], L.epoch + 1);
    }
  }
}

// ---- cp.async (LDGSTS) staging: global -> shared without passing through registers
__device__ __forceinline__ void cp_async16_ca(uint32_t saddr, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(uint32_t saddr, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4_ca(uint32_t saddr, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// shared-window (32-bit) addressed loads
__device__ __forceinline__ uint4 lds16_abs(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(saddr)
               : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds8_abs(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
// per-thread staging ring of the bulk kernels (128 threads per CTA): per stage
// 128 x 16 B opposite-plane vectors, 128 x 16 B own-plane vectors, 128 x 4 B edge words
constexpr int kBulkStages = 4;
constexpr bool kBulkPair = CMG_BULK_PAIR != 0;  // bulk kernels: pair table (1 LDS / 2 sites) or lane tables
constexpr int kSmemRing = kBulkPair ? kSmemTile : 256;
constexpr uint32_t kBulkStageBytes = 128u * 36u;
constexpr int kSmemBulk2d = kSmemRing + kBulkStages * (int)kBulkStageBytes;

// Programmatic dependent launch (the half-sweeps of a run are a chain of kernels on one
// stream): a kernel lets the next one start as soon as its own CTAs are all past their wait,
// so the next half-sweep's CTAs become resident while this one's tail drains, load their
// tables and then wait for this grid (pdl_wait: the whole of it and its memory; chained
// launches: the neighbour CTAs only, hs_wait_for) before they touch the planes.
// Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Chained half-sweeps: instead of pdl_wait (the WHOLE previous grid) a thread waits for the
// CTAs of the previous half-sweep that hold its neighbours -- N of them, ids[] within the
// chain's row of flags -- to have published epoch `want` or later.  The kernels are launched
// as programmatic dependents with the trigger at their start, so every CTA waited for is
// resident or finished (no deadlock), and a CTA only ever waits for CTAs of the half-sweep
// before it.  Relaxed polls, then one acquire fence (which also drops stale L1 lines of the
// plane the previous half-sweep rewrote).  Bounded like every other wait of the library.
__device__ __forceinline__ unsigned int ld_relaxed_gpu_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <int N>
__device__ __forceinline__ void hs_wait_for(const unsigned int *flags, const int (&ids)[N], uint32_t want,
                                            unsigned int *error) {
  unsigned int spins = 0;
  for (;;) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) ok = ok && (int)(ld_relaxed_gpu_u32(flags + ids[i]) - want) >= 0;
    if (ok) break;
    if (++spins > (1u << 22)) {
      if (error) atomicOr(error, kErrChain);
      break;
    }
    __nanosleep(100);
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// after the CTA's last plane store: barrier, then one release store
__device__ __forceinline__ void hs_publish(unsigned int *flag, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}

template <bool SAMPLE, int R = 10>
__global__ void __launch_bounds__(128, CMG_BULK_CTAS) k_halfsweep_bulk2d(SweepArgs A) {
  const LatticeView &L = A.L;
  const int chain = blockIdx.y;
  const int h = L.h, n1 = L.n1;
  const int V = h >> 4;  // 16-byte vectors per column
  // strips: A.n_strips balanced ones (even starts, lengths within two columns of each
  // other) or uniform strips of A.js columns
  const int n_strips = A.n_strips > 0 ? A.n_strips : (n1 + A.js - 1) / A.js;
  const bool edge_mode = L.edge_mode != 0 && n_strips >= 2;
  // the table loads are issued first and only waited for (CTA barrier below)
  // after the thread's pipeline has been filled
  load_accept_table(A.tabs + chain, kBulkPair);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (A.hs_wait) {
    if (t < (long long)V * n_strips) {
      const int v = (int)(t % V), strip = (int)(t / V);
      const int sm = strip == 0 ? n_strips - 1 : strip - 1, sp = strip == n_strips - 1 ? 0 : strip + 1;
      const int vm = v == 0 ? V - 1 : v - 1, vp = v == V - 1 ? 0 : v + 1;
      const int ids[5] = {(int)blockIdx.x, (v + V * sm) >> 7, (v + V * sp) >> 7, (vm + V * strip) >> 7,
                          (vp + V * strip) >> 7};
      hs_wait_for(A.hs_flags + (long long)chain * gridDim.x, ids, A.hs_epoch - 1u, A.error);
    }
  } else {
    pdl_wait();
  }
  // the next half-sweep may become resident once every CTA of this one is past its wait: at
  // most two half-sweeps share the GPU, and everything a CTA waits for has been resident
  pdl_launch_dependents();
  if (!edge_mode) slab_wait_neighbours(L);

  Accum acc = {0u, 0u, 0u, 0u, 0u};
  bool pushed = false;

  {
    // threads past the end of the lattice run the same code with zero columns
    const bool active = t < (long long)V * n_strips;
    const int v = active ? (int)(t % V) : 0;
    int strip = active ? (int)(t / V) : 0;
    // edge mode: the last strip is second in the grid (its CTAs start in the first wave)
    if (edge_mode) strip = strip == 0 ? 0 : (strip == 1 ? n_strips - 1 : strip - 1);
    const int p0 = v << 4;
    int sbeg, send;
    if (A.n_strips > 0) {
      const long long half = n1 >> 1;
      sbeg = 2 * (int)(((long long)strip * half) / n_strips);
      send = 2 * (int)(((long long)(strip + 1) * half) / n_strips);
    } else {
      sbeg = strip * A.js;
      send = min(sbeg + A.js, n1);
    }
    if (!active) send = sbeg;
    uint8_t *C = L.planes + (long long)chain * L.chain_stride +
                 (long long)A.colour * L.plane_stride;
    const uint8_t *O = L.planes + (long long)chain * L.chain_stride +
                       (long long)(1 - A.colour) * L.plane_stride;
    const uint8_t *halo_lo = L.halo_lo[1 - A.colour];
    const uint8_t *halo_hi = L.halo_hi[1 - A.colour];
    uint8_t *push_lo = L.push_lo[A.colour];
    uint8_t *push_hi = L.push_hi[A.colour];
    const uint32_t chain_word = (uint32_t)(chain + A.chain_offset) << 8;

    // column pointer of the opposite plane with periodic wrap / halo
    auto ocol = [&](int j) -> const uint8_t * {
      if (j < 0) return halo_lo ? halo_lo : O + (long long)h * (n1 - 1);
      if (j >= n1) return halo_hi ? halo_hi : O;
      return O + (long long)h * j;
    };

    const int p_below = (p0 == 0) ? h - 1 : p0 - 1;
    const int p_above = (p0 + 16 == h) ? 0 : p0 + 16;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(cmg_smem) + kSmemRing;
    const uint32_t s_o = ring + threadIdx.x * 16u;
    const uint32_t s_c = ring + 128u * 16u + threadIdx.x * 16u;
    const uint32_t s_e = ring + 128u * 32u + threadIdx.x * 4u;
    const unsigned int hstep = (unsigned int)h;
    const int e_lo = p_below & ~3, e_hi = p_above;           // byte 3 / byte 0 of the word
    const unsigned int gstep = (unsigned int)h >> 3;

    // Edge mode: a CTA of the first (last) strip sweeps the two columns at the slab
    // boundary as a segment of their own, ahead of the rest of its strip.  (In edge
    // mode every CTA lies inside one strip, so the segments are CTA-uniform.)
    const int edge_side = !edge_mode ? -1 : (sbeg == 0 ? 0 : (send == n1 ? 1 : -1));
    const int n_seg = (edge_side >= 0 && send - sbeg > 2) ? 2 : 1;
    for (int sg = 0; sg < n_seg; ++sg) {
      int jbeg = sbeg, jend = send;
      if (n_seg == 2) {
        if (edge_side == 0) {
          jbeg = sg == 0 ? 0 : 2;
          jend = sg == 0 ? 2 : send;
        } else {
          jbeg = sg == 0 ? n1 - 2 : sbeg;
          jend = sg == 0 ? n1 : n1 - 2;
        }
      }
      const bool edge_seg = edge_side >= 0 && sg == 0;
      if (edge_seg) slab_wait_side(L, edge_side);
      // Software pipeline through shared memory: every thread owns a private ring of
      // kBulkStages slots and keeps the loads of the next kBulkStages columns in
      // flight with cp.async (LDGSTS), which costs no registers -- with ~270
      // instructions per column and 4-5 warps per scheduler one column of lookahead
      // does not cover an HBM round trip.  Slot k & 3 holds, for pipeline step k:
      // the opposite-plane vector of column jbeg+k+1, the own-plane vector of column
      // jbeg+k and the aligned word holding the byte across the vector edge of
      // column jbeg+k.  The column loop is unrolled by the ring size, so slots and
      // column parities are compile-time constants in the loop body.
      const int n = jend - jbeg;
      const long long jg0 = (long long)jbeg + L.col_offset;
      const int par0 = (int)((jg0 + A.colour) & 1);  // i = 2p + par
      const uint8_t *Oend = ocol(jend) + p0;  // column jend of the opposite plane (wrap / halo)
      const uint8_t *Of = O + (long long)h * (jbeg + 1) + p0;  // fetch front, opposite plane
      const uint8_t *Cf = C + (long long)h * jbeg + p0;        // fetch front, own plane
      const uint8_t *Ef = O + (long long)h * jbeg;             // fetch front, edge words
      auto fetch = [&](const int k, const int par) {
        if (k < n) {
          const uint32_t slot = (uint32_t)(k & (kBulkStages - 1)) * kBulkStageBytes;
          cp_async16_ca(s_o + slot, (k + 1 >= n) ? Oend : Of);
          cp_async16_cg(s_c + slot, Cf);
          cp_async4_ca(s_e + slot, Ef + (par ? e_hi : e_lo));
          Of += hstep;
          Cf += hstep;
          Ef += hstep;
        }
        cp_async_commit();
      };
      uint4 om = make_uint4(0u, 0u, 0u, 0u), oc = om;
      if (n > 0) {
        om = ld16_cg(ocol(jbeg - 1) + p0);  // (not .nc: a chained half-sweep overlaps the one that wrote them)
        oc = ld16_cg(O + (long long)h * jbeg + p0);
      }
#pragma unroll
      for (int k = 0; k < kBulkStages; ++k) fetch(k, par0 ^ (k & 1));
      __syncthreads();  // acceptance tables are in shared memory
      uint8_t *Cp = C + (long long)h * jbeg + p0;  // own plane, column j
      unsigned long long g = (unsigned long long)(((long long)h * jg0 + p0) >> 3);

      auto column = [&](const int it, const int par) {
        const uint32_t slot = (uint32_t)(it & (kBulkStages - 1)) * kBulkStageBytes;
        cp_async_wait<kBulkStages - 1>();
        const uint4 op = lds16_abs(s_o + slot);
        const uint4 ce = lds16_abs(s_c + slot);
        const uint32_t eb = lds8_abs(s_e + slot + (par ? 0u : 3u));
        fetch(it + kBulkStages, par);
        const uint4 side = (par == 0) ? shift_up_1(oc, eb) : shift_down_1(oc, eb);
        const uint4 cn = update16<SAMPLE, kBulkPair, R>(ce, om, oc, op, side, g, A.pass, A.colour, chain_word,
                                             A.rk, acc);
        *reinterpret_cast<uint4 *>(Cp) = cn;
        Cp += hstep;
        g += gstep;
        om = oc;
        oc = op;
      };
      auto strip_loop = [&](auto par_tag) {
        constexpr int P0 = decltype(par_tag)::value;
        static_assert(kBulkStages == 4, "the column loop is unrolled by the ring size");
        int it = 0;
        for (; it + 4 <= n; it += 4) {
          column(it, P0);
          column(it + 1, P0 ^ 1);
          column(it + 2, P0);
          column(it + 3, P0 ^ 1);
        }
        for (; it < n; ++it) column(it, P0 ^ (it & 1));
      };
      if (par0) {
        strip_loop(std::integral_constant<int, 1>{});
      } else {
        strip_loop(std::integral_constant<int, 0>{});
      }
      // slab decomposition: the boundary columns also go to the neighbours' halos
      if (!edge_mode || edge_seg) {
        if (active && push_lo && jbeg == 0) {
          *reinterpret_cast<uint4 *>(push_lo + p0) = ld16(C + p0);
          pushed = true;
        }
        if (active && push_hi && jend == n1) {
          *reinterpret_cast<uint4 *>(push_hi + p0) = ld16(C + (long long)h * (n1 - 1) + p0);
          pushed = true;
        }
      }
      if (edge_seg) slab_signal_side(L, edge_side, (unsigned int)(V >> 7) * gridDim.y);
    }
  }
  long long ones = 0, bsum = 0;
  if (SAMPLE) accum_finish(acc, 4, ones, bsum);
  block_accumulate<128>(acc.acc, ones, bsum, SAMPLE && A.sb, A.n_accept + chain,
                        SAMPLE && A.sb ? A.sb + (long long)chain * A.sb_chain_stride : nullptr);
  if (!edge_mode) slab_signal_neighbours(L, pushed);
  if (A.hs_flags) hs_publish(A.hs_flags + (long long)chain * gridDim.x + blockIdx.x, A.hs_epoch);
}

// ---------------------------------------------------------------------------
// k_tile2d: shared-memory resident 2-d kernel with temporal blocking.
//
// A CTA owns a tile of whole columns (all n0/2 bytes of a plane column, so the
// periodic wrap along i stays inside the tile) of one chain, stages both colour
// planes of the tile plus H = 2*P halo columns on each side in shared memory,
// and advances P passes (2*P half-sweeps) before writing the owned columns
// back.  Half-sweep s can only update local columns [s+1, W-1-s): one more halo
// column goes stale per half-sweep.  The halo work is redundant -- the
// neighbouring tile computes the same sites -- but because the random numbers
// are counter-based (a pure function of site, pass, colour) both CTAs get the
// same answer, so no communication is needed inside the launch.
// With n_tiles == 1 the tile is the whole lattice: no halo, periodic columns,
// any number of passes per launch (the 256x256 chains of the (T, mu) grid).
// Launches per pass drop from 2 to 1/P and the per-site loads become LDS.
// ---------------------------------------------------------------------------
// mbarrier helpers (shared by k_tile2d, k_ring2d, k_halfsweep_tma3d)
__device__ __forceinline__ void mbar_init(unsigned long long *mbar, unsigned int count) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *mbar, unsigned int bytes) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *mbar) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *mbar, unsigned int parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
  unsigned int ok;
  asm volatile(
      "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
      : "=r"(ok)
      : "r"(a), "r"(parity)
      : "memory");
  return ok != 0;
}

struct TileArgs {
  LatticeView L;
  const ChainTables *tabs;
  unsigned long long *n_accept;  // [chain]
  long long *sb;                 // slot of the first sample taken by this launch (chain 0)
  long long sb_chain_stride;     // long long units
  long long sb_slot_stride;      // long long units
  unsigned long long pass0;      // Philox pass index of the first pass
  long long pass_phase;          // passes done before this launch (sample schedule)
  long long sample_period;       // 0 = never
  uint32_t rk[20];
  int n_passes;                  // P
  int n_tiles;
  int halo;                      // 2*P, or 0 when n_tiles == 1
  int w_max;                     // widest tile incl. halos (smem plane = w_max*h bytes)
  uint32_t v_magic;              // ceil(2^32 / V), V = h/16
  int chain_offset;              // global index of chain 0
  unsigned int *error;           // sticky error word (bounded waits)
  // where the owned columns are written back: the planes themselves for one lattice per
  // CTA; the context's second copy of the planes for tiles with halos (the host swaps the two
  // after the launch).  A tile stages its halo columns from global memory, so writing back in
  // place is only right while every tile of a lattice has staged before any finishes -- true
  // for one wave of CTAs on an otherwise idle GPU, false for several waves (16 chains of
  // 1024^2 in 16 waves differed from run to run; found by tools/auto_vs_generic.py).
  uint8_t *out_planes;
};
#ifndef CMG_TILE_CTA_SYNC
#define CMG_TILE_CTA_SYNC 0  // 1: the CTA barrier per half-sweep of the first form (A/B builds)
#endif

template <int NT>
__device__ __forceinline__ void block_add2(long long a, long long b, long long *dst,
                                           long long *sh /* [2*NT/32] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh[warp] = a;
    sh[NT / 32 + warp] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long ta = 0, tb = 0;
    for (int w = 0; w < NT / 32; ++w) {
      ta += sh[w];
      tb += sh[NT / 32 + w];
    }
    atomicAdd((unsigned long long *)&dst[0], (unsigned long long)ta);
    atomicAdd((unsigned long long *)&dst[1], (unsigned long long)tb);
  }
  __syncthreads();
}

// 16-byte shared-memory accesses by byte offset into cmg_smem
__device__ __forceinline__ uint4 lds16(uint32_t off) {
  return *reinterpret_cast<const uint4 *>(cmg_smem + off);
}
__device__ __forceinline__ void sts16(uint32_t off, uint4 v) {
  *reinterpret_cast<uint4 *>(cmg_smem + off) = v;
}

constexpr int kTileMaxPasses = 64;

// SINGLE: the tile is the whole lattice (n_tiles == 1): periodic columns, no halo,
// every column owned -- compiled separately so that the halo bookkeeping is not
// in the loop of the many-small-lattices case.
template <int NT, bool SINGLE>
__global__ void __launch_bounds__(NT, 1) k_tile2d(TileArgs A) {
  __shared__ long long s_acc[2 * kTileMaxPasses];  // per-pass {ones, B} of this CTA
  for (int i = threadIdx.x; i < 2 * kTileMaxPasses; i += NT) s_acc[i] = 0;
  const LatticeView &L = A.L;
  const int chain = blockIdx.y;
  const int tile = blockIdx.x;
  load_accept_table(A.tabs + chain, true);
  // consecutive launches of a run are programmatic dependents: this grid becomes resident
  // and loads its tables under the tail of the one before it
  pdl_wait();
  pdl_launch_dependents();

  const int h = L.h, n1 = L.n1;
  const int V = h >> 4;
  constexpr bool periodic = SINGLE;
  const int H = periodic ? 0 : A.halo;
  const int c0 = (int)(((long long)tile * n1) / A.n_tiles);
  const int c1 = (int)(((long long)(tile + 1) * n1) / A.n_tiles);
  const int TW = c1 - c0;
  const int W = TW + 2 * H;
  // byte offsets of the two colour planes of the tile inside cmg_smem
  const uint32_t soff[2] = {(uint32_t)kSmemTile, (uint32_t)kSmemTile + (uint32_t)A.w_max * (uint32_t)h};
  uint8_t *G[2] = {L.planes + (long long)chain * L.chain_stride,
                   L.planes + (long long)chain * L.chain_stride + L.plane_stride};
  const uint32_t chain_word = (uint32_t)(chain + A.chain_offset) << 8;

  // ---- stage the tile: local column cl <-> global column (c0 - H + cl) mod n1
  for (int it = threadIdx.x; it < 2 * W * V; it += NT) {
    const int plane = it >= W * V;
    const int r = it - plane * W * V;
    const int cl = (int)__umulhi((uint32_t)r, A.v_magic);
    const int v = r - cl * V;
    int gc = c0 - H + cl;
    gc += (gc < 0) ? n1 : 0;
    gc -= (gc >= n1) ? n1 : 0;
    sts16(soff[plane] + (uint32_t)(cl * h + (v << 4)),
          __ldcg(reinterpret_cast<const uint4 *>(G[plane] + (long long)gc * h + (v << 4))));  // (not .nc: the grid may start under its predecessor)
  }
  // One lattice per CTA with a fixed assignment of columns to threads and every column
  // group inside one warp (V <= 32): between half-sweeps a warp waits for the two warps
  // that hold the columns next to its own, not for the CTA (per-warp mbarriers as in
  // k_ring2d; with 256^2 chains a thread has four columns per half-sweep, so a CTA barrier
  // per half-sweep tied sixteen independent warps together every ~1400 cycles).
  __shared__ __align__(8) unsigned long long s_tbar[2][NT / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool fine = !CMG_TILE_CTA_SYNC && SINGLE && NT % V == 0 && V <= 32 && 32 % V == 0 && W >= NT / V;
  if (fine && lane == 0) {
    mbar_init(&s_tbar[0][warp], 2);
    mbar_init(&s_tbar[1][warp], 2);
  }
  const int nbr_warp = lane == 0 ? (warp == 0 ? NT / 32 - 1 : warp - 1) : (warp == NT / 32 - 1 ? 0 : warp + 1);
  __syncthreads();

  unsigned int n_acc = 0;
  int slot = 0;
  for (int s = 0; s < 2 * A.n_passes; ++s) {
    const int colour = s & 1;
    const int pl = s >> 1;
    const unsigned long long pass = A.pass0 + (unsigned long long)pl;
    const int lo = periodic ? 0 : s + 1;
    const int hi = periodic ? W : W - 1 - s;
    const bool sample = colour == 1 && A.sample_period > 0 &&
                        ((A.pass_phase + pl + 1) % A.sample_period) == 0;
    const uint32_t cbase = colour ? soff[1] : soff[0];
    const uint32_t obase = colour ? soff[0] : soff[1];
    // global column of local column lo, its parity and Philox group base
    int gc_lo = c0 - H + lo;
    gc_lo += (gc_lo < 0) ? n1 : 0;
    Accum acc = {0u, 0u, 0u, 0u, 0u};
    // one 16-site vector of column cl (local), vector offset p0 (bytes)
    auto process = [&](int cl, int dc, int p0) {
      int cm = cl - 1, cp = cl + 1;
      if (periodic) {
        cm = (cm < 0) ? W - 1 : cm;
        cp = (cp >= W) ? 0 : cp;
      }
      int gc = gc_lo + dc;
      gc -= (gc >= n1) ? n1 : 0;
      const uint32_t col_off = (uint32_t)(cl * h);
      const uint4 ce = lds16(cbase + col_off + p0);
      const uint4 oc = lds16(obase + col_off + p0);
      const uint4 om = lds16(obase + (uint32_t)(cm * h) + p0);
      const uint4 op = lds16(obase + (uint32_t)(cp * h) + p0);
      uint4 side;
      if (((gc + colour) & 1) == 0) {  // i = 2p: the other i-neighbour is p-1
        side = shift_up_1(oc, cmg_smem[obase + col_off + ((p0 == 0) ? h - 1 : p0 - 1)]);
      } else {  // i = 2p+1: p+1
        side = shift_down_1(oc, cmg_smem[obase + col_off + ((p0 + 16 == h) ? 0 : p0 + 16)]);
      }
      const unsigned long long group0 =
          (unsigned long long)(((long long)h * gc + p0) >> 3);
      const bool owned = periodic || (cl >= H && cl < H + TW);
      uint4 cn;
      if (sample && owned) {
        cn = update16<true, true>(ce, om, oc, op, side, group0, pass, colour, chain_word, A.rk, acc);
      } else {
        Accum scratch = {0u, 0u, 0u, 0u, 0u};
        cn = update16<false, true>(ce, om, oc, op, side, group0, pass, colour, chain_word, A.rk,
                                   scratch);
        if (owned) acc.acc += scratch.acc;
      }
      sts16(cbase + col_off + p0, cn);
    };
    if (NT % V == 0) {
      // Each thread keeps its vector offset and walks a contiguous run of columns
      // with the three opposite-colour columns in a register rolling window (two
      // LDS.128 per 16 sites instead of four), shared-memory offsets and the Philox
      // group advanced incrementally, and the column parity static per copy of the
      // body (columns are taken in pairs).
      const int Q = NT / V, q = threadIdx.x / V;
      const uint32_t p0 = (uint32_t)(threadIdx.x % V) << 4;
      const int ncols = hi - lo, base = ncols / Q, rem = ncols - base * Q;
      int cl = lo + q * base + min(q, rem);
      const int cl1 = cl + base + (q < rem ? 1 : 0);
      if (cl < cl1) {
        int gc = gc_lo + (cl - lo);
        gc -= (gc >= n1) ? n1 : 0;
        const uint32_t gstep = (uint32_t)h >> 3;
        unsigned long long g = (unsigned long long)gstep * (uint32_t)gc + (p0 >> 3);
        const unsigned long long gwrap = (unsigned long long)gstep * (uint32_t)n1;
        uint32_t coff = (uint32_t)(cl * h);  // byte offset of column cl inside a plane
        const uint32_t e_lo = (p0 == 0) ? (uint32_t)h - 1u : p0 - 1u;
        const uint32_t e_hi = (p0 + 16u == (uint32_t)h) ? 0u : p0 + 16u;
        const uint32_t wrap_off = (uint32_t)((W - 1) * h);
        uint4 om = lds16(obase + ((periodic && cl == 0) ? wrap_off : coff - (uint32_t)h) + p0);
        uint4 oc = lds16(obase + coff + p0);
        // compiled once per sampling mode, so a loop body holds one variant of the update
        auto run = [&](auto sample_tag) {
          constexpr bool kSample = decltype(sample_tag)::value;
          auto item = [&](const int par) {
            const uint32_t noff = (periodic && cl == W - 1) ? 0u : coff + (uint32_t)h;
            const uint4 op = lds16(obase + noff + p0);
            const uint4 ce = lds16(cbase + coff + p0);
            const uint32_t eb = cmg_smem[obase + coff + (par ? e_hi : e_lo)];
            const uint4 side = par ? shift_down_1(oc, eb) : shift_up_1(oc, eb);
            const bool owned = periodic || (cl >= H && cl < H + TW);
            uint4 cn;
            if (kSample && owned) {
              cn = update16<true, true>(ce, om, oc, op, side, g, pass, colour, chain_word, A.rk, acc);
            } else {
              Accum scratch = {0u, 0u, 0u, 0u, 0u};
              cn = update16<false, true>(ce, om, oc, op, side, g, pass, colour, chain_word, A.rk,
                                         scratch);
              if (owned) acc.acc += scratch.acc;
            }
            sts16(cbase + coff + p0, cn);
            om = oc;
            oc = op;
            coff += (uint32_t)h;
            g += gstep;
            ++cl;
            if (!SINGLE && ++gc == n1) {  // halo columns past the lattice edge wrap around
              gc = 0;
              g -= gwrap;
            }
          };
          if ((gc + colour) & 1) item(1);  // i = 2p + par; align the pair loop to par = 0
          while (cl + 2 <= cl1) {
            item(0);
            item(1);
          }
          if (cl < cl1) item(0);
        };
        if (sample) {
          run(std::true_type{});
        } else {
          run(std::false_type{});
        }
      }
    } else {
      const int items = (hi - lo) * V;
      for (int it = threadIdx.x; it < items; it += NT) {
        const int dc = (int)__umulhi((uint32_t)it, A.v_magic);
        process(lo + dc, dc, (it - dc * V) << 4);
      }
    }
    n_acc += acc.acc;
    if (sample) {
      // warp partial sums (hardware integer reduction of the raw byte sums) ->
      // per-pass shared accumulators (flushed once at the end)
      Accum wa;
      wa.acc = 0u;
      wa.c1 = __reduce_add_sync(0xffffffffu, acc.c1);
      wa.opp = __reduce_add_sync(0xffffffffu, acc.opp);
      wa.u7 = __reduce_add_sync(0xffffffffu, acc.u7);
      wa.sites = __reduce_add_sync(0xffffffffu, acc.sites);
      if ((threadIdx.x & 31) == 0) {
        long long ones, bsum;
        accum_finish(wa, 4, ones, bsum);
        atomicAdd(reinterpret_cast<unsigned long long *>(&s_acc[2 * slot]), (unsigned long long)ones);
        atomicAdd(reinterpret_cast<unsigned long long *>(&s_acc[2 * slot + 1]), (unsigned long long)bsum);
      }
      ++slot;
    }
    if (fine) {
      __syncwarp();
      if (lane < 2) mbar_arrive(&s_tbar[colour][nbr_warp]);
      unsigned int spins = 0;
      while (!mbar_try_wait(&s_tbar[colour][warp], (unsigned int)pl & 1u)) {
        if (++spins > (1u << 25)) {  // bounded like every wait of the library (try_wait itself sleeps)
          if (A.error) atomicOr(A.error, kErrRingEdge);
          break;
        }
      }
    } else {
      __syncthreads();
    }
  }
  __syncthreads();

  // ---- flush the sampled sums of this launch (one global atomic per slot and quantity)
  for (int i = threadIdx.x; i < 2 * slot; i += NT) {
    long long *dst = A.sb + (long long)(i >> 1) * A.sb_slot_stride +
                     (long long)chain * A.sb_chain_stride + (i & 1);
    atomicAdd(reinterpret_cast<unsigned long long *>(dst), (unsigned long long)s_acc[i]);
  }

  // ---- write the owned columns back
  for (int it = threadIdx.x; it < 2 * TW * V; it += NT) {
    const int plane = it >= TW * V;
    const int r = it - plane * TW * V;
    const int dc = (int)__umulhi((uint32_t)r, A.v_magic);
    const int v = r - dc * V;
    *reinterpret_cast<uint4 *>(A.out_planes + (long long)chain * L.chain_stride + (plane ? L.plane_stride : 0) +
                               (long long)(c0 + dc) * h + (v << 4)) =
        lds16(soff[plane] + (uint32_t)((H + dc) * h + (v << 4)));
  }
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  if ((threadIdx.x & 31) == 0 && n_acc) atomicAdd(A.n_accept + chain, (unsigned long long)n_acc);
}

// ---------------------------------------------------------------------------
// k_ring2d: one lattice resident in the shared memory of the whole GPU.
//
// Every CTA of a cooperative launch (grid <= #SMs, one CTA per SM) owns a
// tile of whole columns for the entire launch and keeps both colour planes of
// it in shared memory; nothing is recomputed.  The only data that crosses a
// tile edge is one boundary column per side and half-sweep.  It travels
// through a mailbox in global memory (L2-resident) whose bytes validate
// themselves: a published byte is  b | stamp << 1  with stamp = 1 + s mod 127
// of the half-sweep s that produced it (occupation bytes only use bit 0), and
// the mailbox is zeroed before the launch.  The consumer thread loads its own
// 16-byte vector with a gpu-coherent load and accepts it when all 16 stamps
// are the expected one -- no flags, no fences, and tearing of the 16-byte
// access would be detected.  A slot is rewritten two half-sweeps later, by a
// thread that has by then consumed the neighbour's next vector, which the
// neighbour produced after reading this one: the dependency chain orders the
// reuse.  The column groups that own an edge start their run at the edge
// (the right-hand group walks its columns downwards), so an edge is published
// one column-time into the half-sweep and consumed at the start of the next;
// the consumer issues its load before the CTA barrier that ends the previous
// half-sweep, so the L2 round trip overlaps the barrier wait.
// Random numbers are counter-based, so the trajectory is the one of every
// other kernel.  Any number of passes per launch (sample slots permitting).
// ---------------------------------------------------------------------------
struct RingArgs {
  LatticeView L;
  const ChainTables *tabs;
  unsigned long long *n_accept;  // [chain]
  long long *sb;                 // slot of the first sample taken by this launch (chain 0)
  long long sb_chain_stride;     // long long units
  long long sb_slot_stride;      // long long units
  unsigned long long pass0;      // Philox pass index of the first pass
  long long pass_phase;          // passes done before this launch (sample schedule)
  long long sample_period;       // 0 = never
  uint32_t rk[20];
  int n_passes;
  int n_tiles;
  int w_max;                     // widest tile (smem plane = w_max*h bytes)
  uint32_t v_magic;              // ceil(2^32 / V), V = h/16
  int chain_offset;              // global index of chain 0
  uint8_t *mailbox;              // [chain][tile][side][plane][h]: what ARRIVES at `side` of `tile`
  unsigned int *error;           // the context's sticky error word (edge wait / bulk copy timed out)
  // A ring that continues on other GPUs (column slabs of one lattice, one cooperative
  // launch per GPU): the outer edge of the first / last tile is published into the
  // mailbox of the low / high neighbour GPU (peer memory over NVLink; null: the ring
  // closes on this GPU) and consumed from our own, with system scope.  The stamps then
  // count the half-sweeps of the whole trajectory (stamp0 = half-sweeps before this
  // launch, mod 127), the mailbox is never zeroed, and the outer edges of the first
  // half-sweep of a launch come from the mailbox too (the last half-sweep of the
  // previous launch, or k_ring_publish, put them there).
  uint8_t *peer_mb[2];
  int peer_tiles[2];             // tiles per chain on that neighbour
  uint32_t stamp0;
};
constexpr int kRingMaxPasses = 256;

__device__ __forceinline__ uint4 ld_relaxed_gpu_v4(const void *p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_v4(void *p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const void *p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_v4(void *p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
// ---- bulk (TMA) copies of a tile: a contiguous run of whole columns of a plane --------
// One thread issues one cp.async.bulk per plane (UBLKCP); completion is counted
// in bytes on an mbarrier that every thread of the CTA then waits for.
// global -> shared, bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void *gsrc, unsigned int bytes,
                                          unsigned long long *mbar) {
  const uint32_t m = (uint32_t)__cvta_generic_to_shared(mbar);
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
      "l"(gsrc), "r"(bytes), "r"(m)
      : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_store(void *gdst, uint32_t smem_src, unsigned int bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_store_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_shared() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ bool ring_stamp_ok(uint4 v, uint32_t expect) {
  return ((((v.x ^ expect) | (v.y ^ expect)) | ((v.z ^ expect) | (v.w ^ expect))) & 0xfefefefeu) == 0u;
}

#ifndef CMG_RING_CTA_SYNC
#define CMG_RING_CTA_SYNC 0
#endif
template <int NT, int R = 10>
__global__ void __launch_bounds__(NT, 1) k_ring2d(RingArgs A) {
  __shared__ long long s_acc[2 * kRingMaxPasses];  // per-pass {ones, B} of this CTA
  for (int i = threadIdx.x; i < 2 * kRingMaxPasses; i += NT) s_acc[i] = 0;
  const LatticeView &L = A.L;
  const int chain = blockIdx.y;
  const int tile = blockIdx.x;
  load_accept_table(A.tabs + chain, true);

  const int h = L.h, n1 = L.n1;
  const int V = h >> 4;
  const int c0 = (int)(((long long)tile * n1) / A.n_tiles);
  const int c1 = (int)(((long long)(tile + 1) * n1) / A.n_tiles);
  const int TW = c1 - c0;
  const uint32_t soff[2] = {(uint32_t)kSmemTile, (uint32_t)kSmemTile + (uint32_t)A.w_max * (uint32_t)h};
  uint8_t *G[2] = {L.planes + (long long)chain * L.chain_stride,
                   L.planes + (long long)chain * L.chain_stride + L.plane_stride};
  const uint32_t chain_word = (uint32_t)(chain + A.chain_offset) << 8;

  // ---- stage the owned columns: the tile of a plane is TW*h contiguous bytes
  // both in global and in shared memory, so it is one bulk (TMA) copy per plane
  __shared__ __align__(8) unsigned long long s_mbar;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(cmg_smem);
  const unsigned int tile_bytes = (unsigned int)(TW * h);
  if (threadIdx.x == 0) mbar_init(&s_mbar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&s_mbar, 2u * tile_bytes);
    bulk_load(smem_base + soff[0], G[0] + (long long)c0 * h, tile_bytes, &s_mbar);
    bulk_load(smem_base + soff[1], G[1] + (long long)c0 * h, tile_bytes, &s_mbar);
  }
  {
    unsigned int spins = 0;
    while (!mbar_try_wait(&s_mbar, 0)) {
      if (++spins > (1u << 12)) {  // bounded (try_wait itself sleeps), like the edge waits
        atomicOr(A.error, kErrRingCopy);
        break;
      }
    }
  }

  // column groups: Q >= 2 groups of V threads; group q owns columns [a, b).
  // Group 0 starts at the left edge and walks up, group Q-1 starts at the
  // right edge and walks down, the others walk up.
  const int Q = NT / V, q = threadIdx.x / V;
  const uint32_t p0 = (uint32_t)(threadIdx.x % V) << 4;
  const int base = TW / Q, rem = TW - base * Q;
  const int a = q * base + min(q, rem);
  const int b = a + base + (q < rem ? 1 : 0);
  const bool down = (q == Q - 1);
  const bool edge = (q == 0) || down;
  const int cl_first = down ? b - 1 : a;
  const int n_cols = b - a;
  const int dh = down ? -h : h;                   // byte step between consecutive columns of the run
  const uint32_t gstep = (uint32_t)h >> 3;
  const long long dg = down ? -(long long)gstep : (long long)gstep;
  const uint32_t e_lo = (p0 == 0) ? (uint32_t)h - 1u : p0 - 1u;
  const uint32_t e_hi = (p0 + 16u == (uint32_t)h) ? 0u : p0 + 16u;
  // mailbox slots, indexed by the CONSUMER: we publish into the slot of the neighbouring
  // tile's side that faces us (on the neighbour GPU for the outer edges of a slab) and
  // consume what arrives at our own side
  int nb = down ? (tile + 1 == A.n_tiles ? 0 : tile + 1) : (tile == 0 ? A.n_tiles - 1 : tile - 1);
  uint8_t *out_base = A.mailbox;
  int out_tiles = A.n_tiles;
  bool remote = false;
  if (down && tile + 1 == A.n_tiles && A.peer_mb[1]) {
    remote = true;
    out_base = A.peer_mb[1];
    out_tiles = A.peer_tiles[1];
    nb = 0;
  } else if (!down && tile == 0 && A.peer_mb[0]) {
    remote = true;
    out_base = A.peer_mb[0];
    out_tiles = A.peer_tiles[0];
    nb = out_tiles - 1;
  }
  const long long slot_bytes = 2ll * h;  // two planes per (tile, side)
  uint8_t *mb_out = out_base + (((long long)chain * out_tiles + nb) * 2 + (down ? 0 : 1)) * slot_bytes + p0;
  const uint8_t *mb_in = A.mailbox + (((long long)chain * A.n_tiles + tile) * 2 + (down ? 1 : 0)) * slot_bytes + p0;
  // the neighbour's edge column at home (first half-sweep of the launch)
  const int gc_nb = down ? (c1 == n1 ? 0 : c1) : (c0 == 0 ? n1 - 1 : c0 - 1);
  uint4 hv = make_uint4(0u, 0u, 0u, 0u);  // (no stamp: an outer edge polls its mailbox at once)
  if (edge && !remote) hv = __ldcg(reinterpret_cast<const uint4 *>(G[1] + (long long)gc_nb * h + p0));
  // Between half-sweeps a warp waits for the warps whose bytes it reads and whose reads it
  // overwrites, not for the CTA: the warps above and below it in its column group (the byte
  // across the vector edge) and the warps of the same rows in the two adjacent groups (the
  // column next to the run).  One mbarrier per warp and half-sweep parity counts the
  // arrivals of those neighbours; a warp is never more than one half-sweep ahead of a
  // neighbour, so two barriers per warp suffice, and the waits are bounded like the others.
  __shared__ __align__(8) unsigned long long s_nbar[2][NT / 32];
  // (column groups narrower than a warp, V < 32 -- the 128-thread form for n0 = 256 / 512: a
  // warp holds several whole groups and borders the warps before and after it only)
  const int wpg = V >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wu = wpg ? warp % wpg : 0;
  int nbr[4], n_nbr = 0;
  {
    const int cand[4] = {wpg ? (q > 0 ? warp - wpg : -1) : (warp > 0 ? warp - 1 : -1),
                         wpg ? (q < Q - 1 ? warp + wpg : -1) : (warp < NT / 32 - 1 ? warp + 1 : -1),
                         wpg ? q * wpg + (wu + 1) % wpg : -1, wpg ? q * wpg + (wu + wpg - 1) % wpg : -1};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bool take = cand[i] >= 0 && cand[i] != warp;
#pragma unroll
      for (int k = 0; k < i; ++k) take = take && cand[k] != cand[i];
      nbr[i] = take ? cand[i] : -1;
      n_nbr += take ? 1 : 0;
    }
  }
  if (lane == 0) {
    mbar_init(&s_nbar[0][warp], (unsigned int)n_nbr);
    mbar_init(&s_nbar[1][warp], (unsigned int)n_nbr);
  }
  const int my_nbr = lane == 0 ? nbr[0] : lane == 1 ? nbr[1] : lane == 2 ? nbr[2] : lane == 3 ? nbr[3] : -1;
  __syncthreads();

  unsigned int n_acc = 0;
  int slot = 0;
  for (int s = 0; s < 2 * A.n_passes; ++s) {
    const int colour = s & 1;
    const int pl = s >> 1;
    const unsigned long long pass = A.pass0 + (unsigned long long)pl;
    const bool sample = colour == 1 && A.sample_period > 0 &&
                        ((A.pass_phase + pl + 1) % A.sample_period) == 0;
    const uint32_t cbase = colour ? soff[1] : soff[0];
    const uint32_t obase = colour ? soff[0] : soff[1];
    const uint32_t stampw = ((A.stamp0 + (uint32_t)s) % 127u + 1u) * 0x02020202u;  // this half-sweep's stamp << 1
    Accum acc = {0u, 0u, 0u, 0u, 0u};

    int cl = cl_first;
    uint32_t coff = (uint32_t)(cl * h);
    unsigned long long g = (unsigned long long)gstep * (unsigned long long)(L.col_offset + c0 + cl) + (p0 >> 3);
    // the column behind the start of the run: the neighbour's edge, or shared memory
    uint4 om;
    if (edge) {
      if (s > 0 || remote) {
        const uint32_t expect = ((A.stamp0 + (uint32_t)s + 126u) % 127u + 1u) * 0x02020202u;
        unsigned int spins = 0;
        while (!ring_stamp_ok(hv, expect)) {
          hv = remote ? ld_relaxed_sys_v4(mb_in + (colour ^ 1) * h) : ld_relaxed_gpu_v4(mb_in + (colour ^ 1) * h);
          if (++spins > (1u << 22)) {  // bounded: report instead of hanging the GPU
            atomicOr(A.error, kErrRingEdge);
            break;
          }
        }
      }
      om = make_uint4(hv.x & 0x01010101u, hv.y & 0x01010101u, hv.z & 0x01010101u, hv.w & 0x01010101u);
    } else {
      om = lds16(obase + coff - (uint32_t)dh + p0);
    }
    uint4 oc = lds16(obase + coff + p0);
    const int cl_pub = edge ? cl_first : -1;
    // the run of this half-sweep, compiled once per sampling mode so that a loop
    // body holds one variant of the update only
    auto run = [&](auto sample_tag) {
      constexpr bool kSample = decltype(sample_tag)::value;
      auto item = [&](const int par) {
        const uint4 op = lds16(obase + coff + (uint32_t)dh + p0);
        const uint4 ce = lds16(cbase + coff + p0);
        const uint32_t eb = cmg_smem[obase + coff + (par ? e_hi : e_lo)];
        const uint4 side = par ? shift_down_1(oc, eb) : shift_up_1(oc, eb);
        const uint4 cn =
            update16<kSample, true, R>(ce, om, oc, op, side, g, pass, colour, chain_word, A.rk, acc);
        sts16(cbase + coff + p0, cn);
        if (cl == cl_pub) {  // our edge column: publish it, stamped
          const uint4 pv = make_uint4(cn.x | stampw, cn.y | stampw, cn.z | stampw, cn.w | stampw);
          if (remote)
            st_relaxed_sys_v4(mb_out + colour * h, pv);
          else
            st_relaxed_gpu_v4(mb_out + colour * h, pv);
        }
        om = oc;
        oc = op;
        coff += (uint32_t)dh;
        g += (unsigned long long)dg;
        cl += down ? -1 : 1;
      };
      int left = n_cols;
      if ((c0 + cl + colour) & 1) {  // align the pair loop to par = 0
        item(1);
        --left;
      }
      while (left >= 2) {
        item(0);
        item(1);
        left -= 2;
      }
      if (left) item(0);
    };
    if (sample) {
      run(std::true_type{});
    } else {
      run(std::false_type{});
    }

    n_acc += acc.acc;
    if (sample) {
      Accum wa;
      wa.acc = 0u;
      wa.c1 = __reduce_add_sync(0xffffffffu, acc.c1);
      wa.opp = __reduce_add_sync(0xffffffffu, acc.opp);
      wa.u7 = __reduce_add_sync(0xffffffffu, acc.u7);
      wa.sites = __reduce_add_sync(0xffffffffu, acc.sites);
      if ((threadIdx.x & 31) == 0) {
        long long ones, bsum;
        accum_finish(wa, 4, ones, bsum);
        atomicAdd(reinterpret_cast<unsigned long long *>(&s_acc[2 * slot]), (unsigned long long)ones);
        atomicAdd(reinterpret_cast<unsigned long long *>(&s_acc[2 * slot + 1]), (unsigned long long)bsum);
      }
      ++slot;
    }
    // the neighbour published its edge of this colour early in this half-sweep:
    // fetch it now, so that the round trip overlaps the barrier
    if (edge) hv = remote ? ld_relaxed_sys_v4(mb_in + colour * h) : ld_relaxed_gpu_v4(mb_in + colour * h);
    if (CMG_RING_CTA_SYNC) {  // the first form: one CTA barrier per half-sweep (kept for A/B builds)
      __syncthreads();
    } else {
      __syncwarp();
      if (my_nbr >= 0) mbar_arrive(&s_nbar[colour][my_nbr]);
      unsigned int spins = 0;
      while (!mbar_try_wait(&s_nbar[colour][warp], (unsigned int)pl & 1u)) {
        if (++spins > (1u << 25)) {  // (try_wait itself sleeps)
          atomicOr(A.error, kErrRingEdge);
          break;
        }
      }
    }
  }
  __syncthreads();

  // ---- flush the sampled sums of this launch
  for (int i = threadIdx.x; i < 2 * slot; i += NT) {
    long long *dst = A.sb + (long long)(i >> 1) * A.sb_slot_stride +
                     (long long)chain * A.sb_chain_stride + (i & 1);
    atomicAdd(reinterpret_cast<unsigned long long *>(dst), (unsigned long long)s_acc[i]);
  }

  // ---- write the owned columns back: one bulk store per plane.  The tile was
  // written with ordinary shared-memory stores, which the async proxy only sees
  // after a proxy fence by the writers and a barrier.
  fence_proxy_async_shared();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(cmg_smem);
    const unsigned int nb = (unsigned int)((c1 - c0) * L.h);
    uint8_t *g0 = L.planes + (long long)blockIdx.y * L.chain_stride + (long long)c0 * L.h;
    bulk_store(g0, sb + (uint32_t)kSmemTile, nb);
    bulk_store(g0 + L.plane_stride, sb + (uint32_t)kSmemTile + (uint32_t)A.w_max * (uint32_t)L.h, nb);
    bulk_store_commit_and_wait();  // shared memory must outlive the reads
  }
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  if ((threadIdx.x & 31) == 0 && n_acc) atomicAdd(A.n_accept + chain, (unsigned long long)n_acc);
}

// Slab ring: before the first launch of a run the outer edge columns of plane 1 (what
// the first half-sweep, colour 0, of the neighbour's outer tile reads) go into the
// neighbours' mailboxes, stamped as the half-sweep before the run.  grid (V / 128 or 1, 2 sides).
__global__ void k_ring_publish(RingArgs A) {
  const LatticeView &L = A.L;
  const int h = L.h, side = blockIdx.y;
  const int p0 = (int)(blockIdx.x * blockDim.x + threadIdx.x) << 4;
  if (p0 >= h || !A.peer_mb[side]) return;
  const uint8_t *src = L.planes + L.plane_stride + (side ? (long long)h * (L.n1 - 1) : 0) + p0;
  const int nb = side ? 0 : A.peer_tiles[0] - 1;
  uint8_t *dst = A.peer_mb[side] + ((long long)nb * 2 + (side ? 0 : 1)) * (2ll * h) + h + p0;  // plane 1 of the slot
  const uint32_t stampw = ((A.stamp0 + 126u) % 127u + 1u) * 0x02020202u;
  const uint4 v = *reinterpret_cast<const uint4 *>(src);
  st_relaxed_sys_v4(dst, make_uint4(v.x | stampw, v.y | stampw, v.z | stampw, v.w | stampw));
}

// Slab ring: after the last launch of a run the four boundary columns (two sides, two
// planes) go into the neighbours' halo buffers, as the streaming kernel leaves them after
// every half-sweep, and the neighbours' flags are raised to epoch + 1: sample_now and the
// streaming kernel find the halos they expect.  grid (ceil(V / 128), 4).
__global__ void k_slab_push_edges(LatticeView L) {
  const int h = L.h, side = blockIdx.y & 1, plane = blockIdx.y >> 1;
  const int p0 = (int)(blockIdx.x * blockDim.x + threadIdx.x) << 4;
  bool pushed = false;
  uint8_t *dst = side ? L.push_hi[plane] : L.push_lo[plane];
  if (p0 < h && dst) {
    const uint8_t *src = L.planes + (long long)plane * L.plane_stride + (side ? (long long)h * (L.n1 - 1) : 0) + p0;
    *reinterpret_cast<uint4 *>(dst + p0) = *reinterpret_cast<const uint4 *>(src);
    pushed = true;
  }
  slab_signal_neighbours(L, pushed);
}

// ---------------------------------------------------------------------------
// k_halfsweep_bulk3d: 3-d simple cubic, n0 % 32 == 0, n1 and n2 even.
// Same scheme; the strip runs along j inside one k-layer, the k+-1 neighbours
// are two extra 16-byte loads per column (served by L2: a k-layer of 512^3 is
// 128 KiB per colour).
// ---------------------------------------------------------------------------
// per stage: 128 x 16 B for each of j+1, k-1, k+1 (opposite plane) and the own
// plane, 128 x 4 B edge words
constexpr uint32_t kBulk3dStageBytes = 128u * 68u;
constexpr int kSmemBulk3d = kSmemRing + kBulkStages * (int)kBulk3dStageBytes;

template <bool SAMPLE>
__global__ void __launch_bounds__(128, CMG_BULK_CTAS) k_halfsweep_bulk3d(SweepArgs A) {
  const LatticeView &L = A.L;
  const int chain = blockIdx.y;
  load_accept_table(A.tabs + chain, kBulkPair);  // waited for after the pipeline fill
  if (!A.hs_wait) pdl_wait();

  const int h = L.h, n1 = L.n1, n2 = L.n2;
  const int V = h >> 4;
  const int n_strips = A.n_strips > 0 ? A.n_strips : (n1 + A.js - 1) / A.js;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  Accum acc = {0u, 0u, 0u, 0u, 0u};

  {
    const bool active = t < (long long)V * n_strips * n2;
    int v, strip, k;
    if (A.pair_layers) {
      // With V < 32 a warp holds 32/V strips.  The column parity of a strip is
      // (jbeg + k + colour) & 1 and selects one of two statically specialised strip
      // loops, so the strips of a warp must agree in it: they are the same strip in
      // layers k, k+2, ... (strip starts are even).
      const int spw = 32 / V;
      const long long w = t >> 5;
      const int g = (int)(threadIdx.x & 31) / V;
      v = (int)(threadIdx.x % V);
      strip = (int)(w % n_strips);
      const long long rest = w / n_strips;
      k = active ? (int)(2 * spw * (rest >> 1) + (rest & 1) + 2 * g) : 0;
    } else {
      v = active ? (int)(t % V) : 0;
      const long long t2 = active ? t / V : 0;
      strip = (int)(t2 % n_strips);
      k = (int)(t2 / n_strips);
    }
    if (A.hs_wait && active) {
      // CTA (within the chain) of the thread that holds vector v2 of unit (strip s2, layer k2)
      auto cta_of = [&](int v2, int s2, int k2) -> int {
        if (A.pair_layers) {
          const int spw = 32 / V;
          const long long rest = 2ll * ((k2 >> 1) / spw) + (k2 & 1);
          return (int)((rest * n_strips + s2) >> 2);
        }
        return (int)(((long long)v2 + (long long)V * (s2 + (long long)n_strips * k2)) >> 7);
      };
      const int sm = strip == 0 ? n_strips - 1 : strip - 1, sp = strip == n_strips - 1 ? 0 : strip + 1;
      const int km = k == 0 ? n2 - 1 : k - 1, kp = k == n2 - 1 ? 0 : k + 1;
      const int vm = v == 0 ? V - 1 : v - 1, vp = v == V - 1 ? 0 : v + 1;
      const int ids[7] = {(int)blockIdx.x,       cta_of(v, sm, k),      cta_of(v, sp, k),     cta_of(v, strip, km),
                          cta_of(v, strip, kp), cta_of(vm, strip, k), cta_of(vp, strip, k)};
      hs_wait_for(A.hs_flags + (long long)chain * gridDim.x, ids, A.hs_epoch - 1u, A.error);
    }
    pdl_launch_dependents();  // (see k_halfsweep_bulk2d)
    const int p0 = v << 4;
    int jbeg, jend;
    if (A.n_strips > 0) {
      const long long half = n1 >> 1;
      jbeg = 2 * (int)(((long long)strip * half) / n_strips);
      jend = 2 * (int)(((long long)(strip + 1) * half) / n_strips);
    } else {
      jbeg = strip * A.js;
      jend = min(jbeg + A.js, n1);
    }
    if (!active) jend = jbeg;
    const long long layer = (long long)h * n1;
    uint8_t *C = L.planes + (long long)chain * L.chain_stride +
                 (long long)A.colour * L.plane_stride + layer * k;
    const uint8_t *Oall = L.planes + (long long)chain * L.chain_stride +
                          (long long)(1 - A.colour) * L.plane_stride;
    const uint8_t *O = Oall + layer * k;
    const uint32_t chain_word = (uint32_t)(chain + A.chain_offset) << 8;

    auto wrapj = [&](int j) { return (j < 0) ? n1 - 1 : (j >= n1 ? 0 : j); };
    const int p_below = (p0 == 0) ? h - 1 : p0 - 1;
    const int p_above = (p0 + 16 == h) ? 0 : p0 + 16;

    // same staging ring as k_halfsweep_bulk2d, with the two k-neighbour vectors
    // of the column added to every stage
    const int n = jend - jbeg;
    const int par0 = (jbeg + k + A.colour) & 1;  // i = 2p + par
    const uint8_t *Oend = O + (long long)h * wrapj(jend) + p0;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(cmg_smem) + kSmemRing;
    const uint32_t s_o = ring + threadIdx.x * 16u;
    const uint32_t s_a = s_o + 128u * 16u;
    const uint32_t s_b = s_o + 128u * 32u;
    const uint32_t s_c = s_o + 128u * 48u;
    const uint32_t s_e = ring + 128u * 64u + threadIdx.x * 4u;
    const unsigned int hstep = (unsigned int)h;
    const long long col0 = (long long)h * jbeg + p0;
    const uint8_t *Of = O + col0 + h;                                              // column j+1
    const uint8_t *Af = Oall + layer * ((k == 0) ? n2 - 1 : k - 1) + col0;         // layer k-1
    const uint8_t *Bf = Oall + layer * ((k == n2 - 1) ? 0 : k + 1) + col0;         // layer k+1
    const uint8_t *Cf = C + col0;
    const uint8_t *Ef = O + (long long)h * jbeg;
    const int e_lo = p_below & ~3, e_hi = p_above;
    auto fetch = [&](const int kk, const int par, const uint32_t slot_idx) {
      if (kk < n) {
        const uint32_t slot = slot_idx * kBulk3dStageBytes;
        cp_async16_ca(s_o + slot, (kk + 1 >= n) ? Oend : Of);
        cp_async16_cg(s_a + slot, Af);
        cp_async16_cg(s_b + slot, Bf);
        cp_async16_cg(s_c + slot, Cf);
        cp_async4_ca(s_e + slot, Ef + (par ? e_hi : e_lo));
        Of += hstep;
        Af += hstep;
        Bf += hstep;
        Cf += hstep;
        Ef += hstep;
      }
      cp_async_commit();
    };
    uint4 om = make_uint4(0u, 0u, 0u, 0u), oc = om;
    if (n > 0) {
      om = ld16_cg(O + (long long)h * wrapj(jbeg - 1) + p0);
      oc = ld16_cg(O + col0);
    }
    // The ring slot and the column parity of a column are functions of its VIRTUAL
    // index iv = it - par0 (it = 0 .. n-1 the column of the strip): slot = iv & 3,
    // par = iv & 1.  A strip that starts on an odd-parity column therefore begins
    // with one peeled column (iv = -1: slot 3, par 1) and then runs the SAME
    // four-column body as every other strip.  Warps of layers k and k+1 (opposite
    // start parities) share one hot loop in the instruction cache instead of two.
    auto fetch_v = [&](const int iv) { fetch(iv + par0, iv & 1, (uint32_t)(iv & (kBulkStages - 1))); };
#pragma unroll
    for (int s = 0; s < kBulkStages; ++s) fetch_v(s - par0);
    __syncthreads();  // acceptance tables are in shared memory
    uint8_t *Cp = C + col0;
    unsigned long long g = (unsigned long long)((layer * k + col0) >> 3);
    const unsigned int gstep = (unsigned int)h >> 3;

    auto column = [&](const int iv, const int par, const uint32_t slot_idx) {
      const uint32_t slot = slot_idx * kBulk3dStageBytes;
      cp_async_wait<kBulkStages - 1>();
      const uint4 op = lds16_abs(s_o + slot);
      const uint4 ka = lds16_abs(s_a + slot);
      const uint4 kb = lds16_abs(s_b + slot);
      const uint4 ce = lds16_abs(s_c + slot);
      const uint32_t eb = lds8_abs(s_e + slot + (par ? 0u : 3u));
      fetch(iv + par0 + kBulkStages, par, slot_idx);
      uint4 side = (par == 0) ? shift_up_1(oc, eb) : shift_down_1(oc, eb);
      // fold the two k-neighbours into the "side" word: sums stay <= 6
      side.x += ka.x + kb.x;
      side.y += ka.y + kb.y;
      side.z += ka.z + kb.z;
      side.w += ka.w + kb.w;
      const uint4 cn = update16<SAMPLE, kBulkPair>(ce, om, oc, op, side, g, A.pass, A.colour,
                                                   chain_word, A.rk, acc);
      *reinterpret_cast<uint4 *>(Cp) = cn;
      Cp += hstep;
      g += gstep;
      om = oc;
      oc = op;
    };
    static_assert(kBulkStages == 4, "the column loop is unrolled by the ring size");
    if (par0 && n > 0) column(-1, 1, 3u);
    const int nv = n - par0;  // columns with virtual index >= 0
    int iv = 0;
    for (; iv + 4 <= nv; iv += 4) {
      column(iv, 0, 0u);
      column(iv + 1, 1, 1u);
      column(iv + 2, 0, 2u);
      column(iv + 3, 1, 3u);
    }
    for (; iv < nv; ++iv) column(iv, iv & 1, (uint32_t)(iv & 3));
  }
  long long ones = 0, bsum = 0;
  if (SAMPLE) accum_finish(acc, 6, ones, bsum);
  block_accumulate<128>(acc.acc, ones, bsum, SAMPLE && A.sb, A.n_accept + chain,
                        SAMPLE && A.sb ? A.sb + (long long)chain * A.sb_chain_stride : nullptr);
  if (A.hs_flags) hs_publish(A.hs_flags + (long long)chain * gridDim.x + blockIdx.x, A.hs_epoch);
}

// ---------------------------------------------------------------------------
// k_halfsweep_tma3d: the 3-d half-sweep with the columns moved by the bulk-copy
// (TMA) engine and shared by the CTA.
//
// A CTA of 128 threads owns K = 128 / V consecutive k-layers (V = 16-byte vectors
// per column) of one strip of columns and walks the strip one column at a time, all
// K layers abreast.  Stage t of its shared-memory ring holds, for column jbeg + t,
// the opposite plane of the K + 2 layers k0-1 .. k0+K and (t >= 1) the own plane of
// the K layers at column jbeg + t - 1: 2K + 2 bulk copies of one whole column each
// (cp.async.bulk, completion counted in bytes on the stage's `full` mbarrier),
// issued by the lanes of one warp in one instruction.  Item i (column jbeg + i)
// reads j+1 and its own vector from stage i + 1 and the two k-neighbours and the
// p-edge byte from stage i, where column j of the layers above and below was put
// for THEIR j+1: an opposite-plane vector crosses L2 -> SM (1 + 2/K) times per
// half-sweep instead of three times (plus the 32-byte sector of the edge word), the
// threads issue no load instructions for the stream, and the L1 data pipe only
// carries the LDS.  A stage is released on its `empty` mbarrier (one arrival per
// warp) when the k-neighbours have been read; warp w refills the stages = w mod 4,
// four stages ahead.  Same update, same random numbers, same trajectory as every
// other kernel.  Host side: V in {16, 32} (n0 = 512 or 1024), n2 % K == 0.
// ---------------------------------------------------------------------------
constexpr int kTmaStages = 8;
__device__ __forceinline__ void mbar_wait_bounded(unsigned long long *mbar, unsigned int parity,
                                                  unsigned int *error) {
  unsigned int spins = 0;
  while (!mbar_try_wait(mbar, parity)) {
    if (++spins > (1u << 22)) {  // bounded (try_wait itself sleeps): report instead of hanging
      if (error) atomicOr(error, kErrRingCopy);
      break;
    }
  }
}
__host__ __device__ inline int tma3d_stage_bytes(int h) {
  const int V = h >> 4, K = 128 / V;
  return (2 * K + 2) * h;
}

template <bool SAMPLE>
__global__ void __launch_bounds__(128, CMG_BULK_CTAS) k_halfsweep_tma3d(SweepArgs A) {
  __shared__ __align__(8) unsigned long long s_full[kTmaStages], s_empty[kTmaStages];
  const LatticeView &L = A.L;
  const int chain = blockIdx.y;
  load_accept_table(A.tabs + chain, true);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTmaStages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 4);
    }
  }
  const int h = L.h, n1 = L.n1, n2 = L.n2;
  const int V = h >> 4, K = 128 / V;
  const int n_strips = A.n_strips;
  const int strip = (int)(blockIdx.x % (unsigned)n_strips);
  const int k0 = (int)(blockIdx.x / (unsigned)n_strips) * K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // layers of a warp share their parity: with V = 16 a warp holds layers w and w + 4
  int kk, v;
  if (V <= 32) {
    v = lane % V;
    kk = warp + 4 * (lane / V);
  } else {
    const int wpl = V >> 5;
    kk = warp / wpl;
    v = (warp % wpl) * 32 + lane;
  }
  const int k = k0 + kk;
  const uint32_t p0 = (uint32_t)v << 4;
  const long long half = n1 >> 1;
  const int jbeg = 2 * (int)(((long long)strip * half) / n_strips);
  const int jend = 2 * (int)(((long long)(strip + 1) * half) / n_strips);
  const int n = jend - jbeg;
  const long long layer = (long long)h * n1;
  uint8_t *Cplane = L.planes + (long long)chain * L.chain_stride + (long long)A.colour * L.plane_stride;
  const uint8_t *Oall = L.planes + (long long)chain * L.chain_stride + (long long)(1 - A.colour) * L.plane_stride;
  const uint32_t chain_word = (uint32_t)(chain + A.chain_offset) << 8;
  const uint32_t SB = (uint32_t)((2 * K + 2) * h);
  const uint32_t opp_bytes = (uint32_t)((K + 2) * h);
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(cmg_smem) + kSmemTile;

  // ---- the loader role (a whole warp; lane l < K+2 copies opposite layer k0-1+l,
  // lane K+2+m copies own layer k0+m)
  const uint8_t *src = nullptr;
  uint32_t dst = 0;
  if (lane < K + 2) {
    int kl = k0 - 1 + lane;
    kl = kl < 0 ? n2 - 1 : (kl >= n2 ? kl - n2 : kl);
    src = Oall + layer * kl + (long long)h * jbeg;
    dst = ring + (uint32_t)lane * (uint32_t)h;
  } else if (lane < 2 * K + 2) {
    src = Cplane + layer * (k0 + lane - (K + 2)) + (long long)h * (jbeg - 1);
    dst = ring + opp_bytes + (uint32_t)(lane - (K + 2)) * (uint32_t)h;
  }
  auto load_stage = [&](const int t) {
    if (t > n) return;
    const int s = t & (kTmaStages - 1), u = t / kTmaStages;
    if (u >= 1) mbar_wait_bounded(&s_empty[s], (unsigned)(u - 1) & 1u, L.error);
    if (lane == 0) mbar_expect_tx(&s_full[s], t >= 1 ? SB : opp_bytes);
    __syncwarp();
    if (lane < K + 2) {
      // the column after the last one of the lattice is column 0
      const long long off = (jbeg + t == n1) ? -(long long)h * jbeg : (long long)h * t;
      bulk_load(dst + (uint32_t)s * SB, src + off, (unsigned)h, &s_full[s]);
    } else if (lane < 2 * K + 2 && t >= 1) {
      bulk_load(dst + (uint32_t)s * SB, src + (long long)h * t, (unsigned)h, &s_full[s]);
    }
  };
  __syncthreads();  // mbarriers initialised, tables in shared memory
  if (warp < 3) load_stage(warp);

  Accum acc = {0u, 0u, 0u, 0u, 0u};
  {
    const uint8_t *O = Oall + layer * k;
    const int par0 = (jbeg + k + A.colour) & 1;  // i = 2p + par
    // thread-constant shared-memory offsets inside a stage
    const uint32_t a_ka = ring + (uint32_t)kk * (uint32_t)h + p0;            // opposite layer k-1
    const uint32_t a_oc = a_ka + (uint32_t)h;                                // opposite layer k
    const uint32_t a_kb = a_oc + (uint32_t)h;                                // opposite layer k+1
    const uint32_t a_ce = ring + opp_bytes + (uint32_t)kk * (uint32_t)h + p0;  // own layer k
    const uint32_t row = ring + (uint32_t)(kk + 1) * (uint32_t)h;
    const uint32_t a_e0 = row + ((p0 == 0) ? (uint32_t)h - 1u : p0 - 1u);
    const uint32_t a_e1 = row + ((p0 + 16u == (uint32_t)h) ? 0u : p0 + 16u);
    uint4 om = make_uint4(0u, 0u, 0u, 0u), oc = om;
    if (n > 0) {
      om = ld16_nc(O + (long long)h * (jbeg == 0 ? n1 - 1 : jbeg - 1) + p0);
      mbar_wait_bounded(&s_full[0], 0u, L.error);
      oc = lds16_abs(a_oc);
    }
    uint8_t *Cp = Cplane + layer * k + (long long)h * jbeg + p0;
    unsigned long long g = (unsigned long long)((layer * k + (long long)h * jbeg + p0) >> 3);
    const unsigned int gstep = (unsigned int)h >> 3;
    const unsigned int hstep = (unsigned int)h;

    auto item = [&](const int i, const int par) {
      const uint32_t s1 = (uint32_t)((i + 1) & (kTmaStages - 1)), s0 = (uint32_t)(i & (kTmaStages - 1));
      const uint32_t o1 = s1 * SB, o0 = s0 * SB;
      mbar_wait_bounded(&s_full[s1], ((unsigned)(i + 1) / kTmaStages) & 1u, L.error);
      const uint4 op = lds16_abs(a_oc + o1);
      const uint4 ce = lds16_abs(a_ce + o1);
      const uint4 ka = lds16_abs(a_ka + o0);
      const uint4 kb = lds16_abs(a_kb + o0);
      const uint32_t eb = lds8_abs((par ? a_e1 : a_e0) + o0);
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s0]);  // stage i: every read of this warp is done
      if (warp == (i & 3)) load_stage(i + 3);
      uint4 side = (par == 0) ? shift_up_1(oc, eb) : shift_down_1(oc, eb);
      side.x += ka.x + kb.x;
      side.y += ka.y + kb.y;
      side.z += ka.z + kb.z;
      side.w += ka.w + kb.w;
      const uint4 cn = update16<SAMPLE, true>(ce, om, oc, op, side, g, A.pass, A.colour, chain_word,
                                              A.rk, acc);
      *reinterpret_cast<uint4 *>(Cp) = cn;
      Cp += hstep;
      g += gstep;
      om = oc;
      oc = op;
    };
    // one hot four-column body for both start parities (see k_halfsweep_bulk3d)
    int i = 0;
    if (par0 && n > 0) {
      item(0, 1);
      i = 1;
    }
    for (; i + 4 <= n; i += 4) {
      item(i, 0);
      item(i + 1, 1);
      item(i + 2, 0);
      item(i + 3, 1);
    }
    for (; i < n; ++i) item(i, (i + par0) & 1);
  }
  long long ones = 0, bsum = 0;
  if (SAMPLE) accum_finish(acc, 6, ones, bsum);
  block_accumulate<128>(acc.acc, ones, bsum, SAMPLE, A.n_accept + chain,
                        SAMPLE ? A.sb + (long long)chain * A.sb_chain_stride : nullptr);
}

// ---------------------------------------------------------------------------
// Natural-order helpers.  "Natural" = one int8 b per site in the reference's
// linear order l = i + n0*(j + n1*k) (model.hh:82-99); used at the API edge,
// by the probes and by the serial reference mode, and as the only layout for
// lattices with an odd extent (which cannot be two-coloured periodically).
// ---------------------------------------------------------------------------
struct NaturalShape {
  int n0, n1, n2, dim;
  long long n_sites;
};

__device__ __forceinline__ long long plane_addr(const NaturalShape &s, long long l,
                                                long long plane_stride) {
  const int i = (int)(l % s.n0);
  const long long r = l / s.n0;
  const int j = (int)(r % s.n1);
  const int k = (int)(r / s.n1);
  const int c = (i + j + k) & 1;
  return (long long)c * plane_stride + (i >> 1) +
         (long long)(s.n0 >> 1) * (j + (long long)s.n1 * k);
}

// int32 (+1/-1, natural) -> planes ; flags any value that is not +-1
__global__ void k_i32_to_planes(const int32_t *__restrict__ src, uint8_t *planes,
                                long long plane_stride, NaturalShape s, int *bad) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  const int v = src[l];
  if (v != 1 && v != -1) atomicExch(bad, 1);
  planes[plane_addr(s, l, plane_stride)] = (uint8_t)(v > 0);
}
__global__ void k_planes_to_i32(const uint8_t *__restrict__ planes,
                                long long plane_stride, int32_t *dst, NaturalShape s) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  dst[l] = planes[plane_addr(s, l, plane_stride)] ? 1 : -1;
}
__global__ void k_i32_to_natural(const int32_t *__restrict__ src, uint8_t *nat,
                                 long long n, int *bad) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n) return;
  const int v = src[l];
  if (v != 1 && v != -1) atomicExch(bad, 1);
  nat[l] = (uint8_t)(v > 0);
}
__global__ void k_natural_to_i32(const uint8_t *__restrict__ nat, int32_t *dst,
                                 long long n) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n) return;
  dst[l] = nat[l] ? 1 : -1;
}
__global__ void k_planes_to_natural(const uint8_t *__restrict__ planes,
                                    long long plane_stride, uint8_t *nat,
                                    NaturalShape s) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  nat[l] = planes[plane_addr(s, l, plane_stride)];
}
__global__ void k_natural_to_planes(const uint8_t *__restrict__ nat, uint8_t *planes,
                                    long long plane_stride, NaturalShape s) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  planes[plane_addr(s, l, plane_stride)] = nat[l];
}
// Compact host formats of the occupation (same site order l as the int32 form):
// int8 +1/-1 per site, and one bit per site (bit l & 7 of byte l >> 3 set iff
// s_l = +1).  `base` is the chain's planes (planar != 0) or its natural array.
__device__ __forceinline__ long long site_addr_of(const NaturalShape &s, long long l, int planar,
                                                  long long plane_stride) {
  return planar ? plane_addr(s, l, plane_stride) : l;
}
__global__ void k_i8_to_sites(const int8_t *__restrict__ src, uint8_t *base, long long plane_stride,
                              NaturalShape s, int planar, int *bad) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  const int v = src[l];
  if (v != 1 && v != -1) atomicExch(bad, 1);
  base[site_addr_of(s, l, planar, plane_stride)] = (uint8_t)(v > 0);
}
__global__ void k_sites_to_i8(const uint8_t *__restrict__ base, long long plane_stride, int8_t *dst,
                              NaturalShape s, int planar) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  dst[l] = base[site_addr_of(s, l, planar, plane_stride)] ? 1 : -1;
}
__global__ void k_bits_to_sites(const uint8_t *__restrict__ bits, uint8_t *base,
                                long long plane_stride, NaturalShape s, int planar) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  base[site_addr_of(s, l, planar, plane_stride)] = (uint8_t)((bits[l >> 3] >> (l & 7)) & 1u);
}
// one thread per output byte (8 sites)
__global__ void k_sites_to_bits(const uint8_t *__restrict__ base, long long plane_stride,
                                uint8_t *bits, NaturalShape s, int planar) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (8 * g >= s.n_sites) return;
  unsigned int v = 0;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const long long l = 8 * g + b;
    if (l < s.n_sites) v |= (unsigned int)(base[site_addr_of(s, l, planar, plane_stride)] & 1u) << b;
  }
  bits[g] = (uint8_t)v;
}
__global__ void k_fill_bytes(uint8_t *dst, long long n, uint8_t v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}
// synthetic input: i.i.d. b with P(b=1) = p_up, keyed on the natural index
// g_offset: index of the context's first group of four sites in the GLOBAL
// lattice (a slab starts at column col_begin), so a decomposed lattice draws the
// same initial state as the undecomposed one
__global__ void k_randomize_natural(uint8_t *nat, long long n, unsigned long long seed,
                                    uint32_t thr_m1, int always, long long g_offset) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (4 * g >= n) return;
  const long long gg = g + g_offset;
  uint32_t rk[20];
  for (int i = 0; i < 10; ++i) {
    rk[2 * i] = (uint32_t)seed + i * kPhiloxW0;
    rk[2 * i + 1] = (uint32_t)(seed >> 32) + i * kPhiloxW1;
  }
  const uint4 r = philox4x32<10>(make_uint4((uint32_t)gg, (uint32_t)(gg >> 32), 0x5EEDu, 0u), rk);
  const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
  for (int w = 0; w < 4; ++w)
    if (4 * g + w < n) nat[4 * g + w] = (uint8_t)(always || rr[w] <= thr_m1);
}

// neighbour count and own occupation of site l, natural layout
__device__ __forceinline__ int natural_n_up(const uint8_t *nat, const NaturalShape &s,
                                            long long l) {
  const int i = (int)(l % s.n0);
  const long long r = l / s.n0;
  const int j = (int)(r % s.n1);
  const int k = (int)(r / s.n1);
  const long long base = (long long)s.n0 * (j + (long long)s.n1 * k);
  const int ip = (i + 1 == s.n0) ? 0 : i + 1, im = (i == 0) ? s.n0 - 1 : i - 1;
  const int jp = (j + 1 == s.n1) ? 0 : j + 1, jm = (j == 0) ? s.n1 - 1 : j - 1;
  const long long kk = (long long)s.n1 * k;
  int n = nat[base + ip] + nat[base + im] + nat[i + (long long)s.n0 * (jp + kk)] +
          nat[i + (long long)s.n0 * (jm + kk)];
  if (s.dim == 3) {
    const int kp = (k + 1 == s.n2) ? 0 : k + 1, km = (k == 0) ? s.n2 - 1 : k - 1;
    n += nat[i + (long long)s.n0 * (j + (long long)s.n1 * kp)] +
         nat[i + (long long)s.n0 * (j + (long long)s.n1 * km)];
  }
  return n;
}

// Integer observables from the natural layout:
// ones = #(+1), B = sum_l s_l*(s_{+i} + s_{+j} [+ s_{+k}])   (model.hh:266-270)
__global__ void __launch_bounds__(256) k_observables_natural(const uint8_t *__restrict__ nat,
                                                             NaturalShape s,
                                                             long long *out /* {ones,B} */) {
  long long ones = 0, bsum = 0;
  for (long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x; l < s.n_sites;
       l += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(l % s.n0);
    const long long r = l / s.n0;
    const int j = (int)(r % s.n1);
    const int k = (int)(r / s.n1);
    const int b = nat[l];
    const int ip = (i + 1 == s.n0) ? 0 : i + 1;
    const int jp = (j + 1 == s.n1) ? 0 : j + 1;
    int nb = (2 * nat[ip + (long long)s.n0 * (j + (long long)s.n1 * k)] - 1) +
             (2 * nat[i + (long long)s.n0 * (jp + (long long)s.n1 * k)] - 1);
    if (s.dim == 3) {
      const int kp = (k + 1 == s.n2) ? 0 : k + 1;
      nb += 2 * nat[i + (long long)s.n0 * (j + (long long)s.n1 * kp)] - 1;
    }
    ones += b;
    bsum += (2 * b - 1) * nb;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ones += __shfl_xor_sync(0xffffffffu, ones, o);
    bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
  }
  __shared__ long long sh[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh[0][warp] = ones;
    sh[1][warp] = bsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long a = 0, b = 0;
    for (int w = 0; w < 8; ++w) {
      a += sh[0][w];
      b += sh[1][w];
    }
    atomicAdd((unsigned long long *)&out[0], (unsigned long long)a);
    atomicAdd((unsigned long long *)&out[1], (unsigned long long)b);
  }
}

// Integer observables straight from the colour planes (even extents): one
// thread per 4 plane indices of colour 1.
__global__ void __launch_bounds__(256) k_observables_planes(LatticeView L, int chain,
                                                            long long *out) {
  const uint8_t *C = L.planes + (long long)chain * L.chain_stride + L.plane_stride;
  const uint8_t *O = L.planes + (long long)chain * L.chain_stride;
  const long long plane_size = (long long)L.h * L.n1 * L.n2;
  const int z = 2 * L.dim;
  long long ones = 0, bsum = 0;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < plane_size;
       q += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(q % L.h);
    const long long jk = q / L.h;
    const int j = (int)(jk % L.n1);
    const int k = (int)(jk / L.n1);
    const int par = (j + k + 1) & 1;
    const int jm = (j == 0) ? L.n1 - 1 : j - 1, jp = (j == L.n1 - 1) ? 0 : j + 1;
    const long long rowk = (long long)L.n1 * k;
    const uint8_t *lo = (L.halo_lo[0] && j == 0) ? L.halo_lo[0] + p : O + p + (long long)L.h * (jm + rowk);
    const uint8_t *hi = (L.halo_hi[0] && j == L.n1 - 1) ? L.halo_hi[0] + p : O + p + (long long)L.h * (jp + rowk);
    int n_up = *lo + *hi + O[p + (long long)L.h * (j + rowk)];
    const int ps = par ? ((p == L.h - 1) ? 0 : p + 1) : ((p == 0) ? L.h - 1 : p - 1);
    n_up += O[ps + (long long)L.h * (j + rowk)];
    if (L.dim == 3) {
      const int km = (k == 0) ? L.n2 - 1 : k - 1, kp = (k == L.n2 - 1) ? 0 : k + 1;
      n_up += O[p + (long long)L.h * (j + (long long)L.n1 * km)] +
              O[p + (long long)L.h * (j + (long long)L.n1 * kp)];
    }
    const int b = C[q];
    ones += b + O[q];
    bsum += (2 * b - 1) * (2 * n_up - z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ones += __shfl_xor_sync(0xffffffffu, ones, o);
    bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
  }
  __shared__ long long sh[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh[0][warp] = ones;
    sh[1][warp] = bsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long a = 0, b = 0;
    for (int w = 0; w < 8; ++w) {
      a += sh[0][w];
      b += sh[1][w];
    }
    atomicAdd((unsigned long long *)&out[0], (unsigned long long)a);
    atomicAdd((unsigned long long *)&out[1], (unsigned long long)b);
  }
}

// use_nlist=false energy form (model.hh:273-285): integer dot products of each
// "row" i with row i+1 and of each "column" j with column j+1 (2-d, natural).
__global__ void k_line_dots(const uint8_t *__restrict__ nat, NaturalShape s,
                            long long *row_dots, long long *col_dots) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  const int i = (int)(l % s.n0);
  const int j = (int)(l / s.n0);
  const int sv = 2 * nat[l] - 1;
  const int ip = (i + 1 == s.n0) ? 0 : i + 1;
  const int jp = (j + 1 == s.n1) ? 0 : j + 1;
  const int right = 2 * nat[ip + (long long)s.n0 * j] - 1;
  const int down = 2 * nat[i + (long long)s.n0 * jp] - 1;
  atomicAdd((unsigned long long *)&row_dots[i], (unsigned long long)(long long)(sv * right));
  atomicAdd((unsigned long long *)&col_dots[j], (unsigned long long)(long long)(sv * down));
}

// ---------------------------------------------------------------------------
// Parity probes (natural layout).  dE_per_site[l] = table dE of flipping l;
// accept[l] = (dE < 0) || (u[l] < exp(-dE*beta))   (methods/metropolis.hh:28-34)
// ---------------------------------------------------------------------------
__global__ void k_delta_e_probe(const uint8_t *__restrict__ nat, NaturalShape s,
                                const ChainTables *tab, double *dE) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  dE[l] = tab->dE[2 * natural_n_up(nat, s, l) + nat[l]];
}
__global__ void k_accept_probe(const uint8_t *__restrict__ nat, NaturalShape s,
                               const ChainTables *tab, const double *__restrict__ u,
                               uint8_t *accept) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  const int idx = 2 * natural_n_up(nat, s, l) + nat[l];
  const double d = tab->dE[idx];
  bool a = d < 0.0;
  if (!a) a = u[l] < tab->prob[idx];
  accept[l] = a ? 1 : 0;
}

// ---------------------------------------------------------------------------
// Single-site access and event deltas (the calculator interface used one event
// at a time, e.g. by a host-driven loop).  `base` is either the natural array
// or the planes of the chain; planar != 0 selects plane addressing.
// ---------------------------------------------------------------------------
__device__ __forceinline__ long long site_addr(const NaturalShape &s, long long l, int planar,
                                               long long plane_stride) {
  return planar ? plane_addr(s, l, plane_stride) : l;
}
__device__ __forceinline__ long long wrap_site(const NaturalShape &s, int i, int j, int k) {
  return i + (long long)s.n0 * (j + (long long)s.n1 * k);
}
__global__ void k_get_occ(const uint8_t *base, NaturalShape s, int planar, long long plane_stride,
                          long long l, int *out) {
  *out = base[site_addr(s, l, planar, plane_stride)] ? 1 : -1;
}
__global__ void k_set_occ(uint8_t *base, NaturalShape s, int planar, long long plane_stride,
                          long long l, int value) {
  base[site_addr(s, l, planar, plane_stride)] = (uint8_t)(value > 0);
}
// model.hh:354-379 and :425-435, one thread.  n_event <= kMaxEventSites.
constexpr int kMaxEventSites = 64;
__global__ void k_event_delta(uint8_t *base, NaturalShape s, int planar, long long plane_stride,
                              double J, int n_event, const long long *ls, const int *new_occ,
                              double *out /* {dE_f, dNx} */) {
  double dE = 0.0, dNx = 0.0;
  uint8_t orig[kMaxEventSites];
  for (int e = 0; e < n_event; ++e) {
    const long long l = ls[e];
    const int i = (int)(l % s.n0);
    const long long r = l / s.n0;
    const int j = (int)(r % s.n1);
    const int k = (int)(r / s.n1);
    const int ip = (i + 1 == s.n0) ? 0 : i + 1, im = (i == 0) ? s.n0 - 1 : i - 1;
    const int jp = (j + 1 == s.n1) ? 0 : j + 1, jm = (j == 0) ? s.n1 - 1 : j - 1;
    int nb = 0;
    nb += base[site_addr(s, wrap_site(s, ip, j, k), planar, plane_stride)] ? 1 : -1;
    nb += base[site_addr(s, wrap_site(s, im, j, k), planar, plane_stride)] ? 1 : -1;
    nb += base[site_addr(s, wrap_site(s, i, jp, k), planar, plane_stride)] ? 1 : -1;
    nb += base[site_addr(s, wrap_site(s, i, jm, k), planar, plane_stride)] ? 1 : -1;
    if (s.dim == 3) {
      const int kp = (k + 1 == s.n2) ? 0 : k + 1, km = (k == 0) ? s.n2 - 1 : k - 1;
      nb += base[site_addr(s, wrap_site(s, i, j, kp), planar, plane_stride)] ? 1 : -1;
      nb += base[site_addr(s, wrap_site(s, i, j, km), planar, plane_stride)] ? 1 : -1;
    }
    const long long a = site_addr(s, l, planar, plane_stride);
    orig[e] = base[a];
    const int ds = new_occ[e] - (orig[e] ? 1 : -1);
    dE = __dadd_rn(dE, __dmul_rn(__dmul_rn(-J, (double)ds), (double)nb));
    dNx = __dadd_rn(dNx, __ddiv_rn((double)ds, 2.0));
    if (n_event > 1) base[a] = (uint8_t)(new_occ[e] > 0);  // applied while summing (:361-371)
  }
  if (n_event > 1)  // un-applied in forward order, as the reference does (:374-376)
    for (int e = 0; e < n_event; ++e)
      base[site_addr(s, ls[e], planar, plane_stride)] = orig[e];
  out[0] = dE;
  out[1] = dNx;
}

// ---------------------------------------------------------------------------
// Serial reference mode.  One thread per chain walks the reference's loop
// (methods/basic_occupation_metropolis.hh:381-404) on the reference's random
// stream: std::mt19937_64, libstdc++-13 uniform_int_distribution<long>(0,N-1)
// (Lemire's nearly-divisionless method on 64-bit words) for the site, and
// generate_canonical<double,53> for the acceptance draw, drawn only when
// dE >= 0.  S and B are tracked incrementally (exact integers) and written to
// the sample series at the sampled pass boundaries.
// ---------------------------------------------------------------------------
struct MT64State {
  unsigned long long x[312];
  int pos;
  int pad;
};

__device__ __forceinline__ void mt64_twist(unsigned long long *x) {
  constexpr unsigned long long UPPER = 0xFFFFFFFF80000000ull, LOWER = 0x7FFFFFFFull;
  constexpr unsigned long long A = 0xB5026F5AA96619E9ull;
  for (int k = 0; k < 312 - 156; ++k) {
    unsigned long long y = (x[k] & UPPER) | (x[k + 1] & LOWER);
    x[k] = x[k + 156] ^ (y >> 1) ^ ((y & 1ull) ? A : 0ull);
  }
  for (int k = 312 - 156; k < 311; ++k) {
    unsigned long long y = (x[k] & UPPER) | (x[k + 1] & LOWER);
    x[k] = x[k + (156 - 312)] ^ (y >> 1) ^ ((y & 1ull) ? A : 0ull);
  }
  unsigned long long y = (x[311] & UPPER) | (x[0] & LOWER);
  x[311] = x[155] ^ (y >> 1) ^ ((y & 1ull) ? A : 0ull);
}
__device__ __forceinline__ unsigned long long mt64_next(unsigned long long *x, int &pos) {
  if (pos >= 312) {
    mt64_twist(x);
    pos = 0;
  }
  unsigned long long z = x[pos++];
  z ^= (z >> 29) & 0x5555555555555555ull;
  z ^= (z << 17) & 0x71D67FFFEDA60000ull;
  z ^= (z << 37) & 0xFFF7EEE000000000ull;
  z ^= (z >> 43);
  return z;
}
// libstdc++ 13 bits/uniform_int_dist.h:252-281 (_S_nd with 128-bit products),
// range = maximum_value + 1 (must be < 2^64)
__device__ __forceinline__ unsigned long long mt64_uniform_int(unsigned long long *x,
                                                               int &pos,
                                                               unsigned long long range) {
  unsigned long long u = mt64_next(x, pos);
  unsigned long long low = u * range;
  unsigned long long high = __umul64hi(u, range);
  if (low < range) {
    const unsigned long long threshold = (0ull - range) % range;
    while (low < threshold) {
      u = mt64_next(x, pos);
      low = u * range;
      high = __umul64hi(u, range);
    }
  }
  return high;
}
// libstdc++ 13 bits/random.tcc:3349-3381 (one 64-bit draw for 53 bits), then
// uniform_real_distribution: r*(b-a)+a  (bits/random.h:1904-1910)
__device__ __forceinline__ double mt64_uniform_real(unsigned long long *x, int &pos,
                                                    double maximum_value) {
  const unsigned long long u = mt64_next(x, pos);
  // u / 2^64: scaling by a power of two is exact, so the product with 2^-64 is
  // the correctly rounded quotient (no software division in the loop)
  double r = __dmul_rn(__ull2double_rn(u), 5.42101086242752217003726400434970855712890625e-20);
  if (r >= 1.0) r = 0.99999999999999988897769753748434595763683319091796875;
  return __dadd_rn(__dmul_rn(r, __dsub_rn(maximum_value, 0.0)), 0.0);
}

struct SerialArgs {
  uint8_t *nat;             // [chain][n_sites]
  NaturalShape shape;
  const ChainTables *tabs;
  MT64State *engines;       // [chain]
  unsigned long long *n_accept;  // [chain]
  long long *cur_sb;        // [chain][2] current {ones, B}, updated in place
  long long *series;        // chain 0 sample slot base {ones,B}; may be null
  long long series_chain_stride;
  long long n_passes;
  long long sample_period;  // 0 = never
  long long pass_base;      // passes already done (for the sample schedule)
  int use_smem;             // lattice staged in shared memory
};

// neighbour count of site l with 32-bit index arithmetic (n_sites < 2^31; the 64-bit
// divisions of natural_n_up are software routines)
__device__ __forceinline__ int natural_n_up32(const uint8_t *nat, int n0, int n1, int n2, int dim,
                                              unsigned int l) {
  const unsigned int r = l / (unsigned int)n0;
  const int i = (int)(l - r * (unsigned int)n0);
  int j = (int)r, k = 0;
  if (dim == 3) {
    k = (int)(r / (unsigned int)n1);
    j = (int)(r - (unsigned int)k * (unsigned int)n1);
  }
  const unsigned int base = (unsigned int)n0 * ((unsigned int)j + (unsigned int)n1 * (unsigned int)k);
  const int ip = (i + 1 == n0) ? 0 : i + 1, im = (i == 0) ? n0 - 1 : i - 1;
  const int jp = (j + 1 == n1) ? 0 : j + 1, jm = (j == 0) ? n1 - 1 : j - 1;
  const unsigned int kk = (unsigned int)n1 * (unsigned int)k;
  int n = nat[base + ip] + nat[base + im] + nat[i + (unsigned int)n0 * ((unsigned int)jp + kk)] +
          nat[i + (unsigned int)n0 * ((unsigned int)jm + kk)];
  if (dim == 3) {
    const int kp = (k + 1 == n2) ? 0 : k + 1, km = (k == 0) ? n2 - 1 : k - 1;
    n += nat[i + (unsigned int)n0 * ((unsigned int)j + (unsigned int)n1 * (unsigned int)kp)] +
         nat[i + (unsigned int)n0 * ((unsigned int)j + (unsigned int)n1 * (unsigned int)km)];
  }
  return n;
}

// One block per chain.  The random stream is produced in blocks of 312 words by
// the whole CTA (the Mersenne twist and the tempering are data-parallel within a
// block of the recurrence), the Metropolis loop itself is walked by thread 0 in the
// reference's order; it pauses whenever the block of words is used up, wherever in
// a step that happens (a step draws one word for the site, redraws with the tiny
// probability of Lemire's rejection, and one more word only when dE >= 0).
constexpr int kSerialThreads = 128;
__device__ __forceinline__ unsigned long long mt64_temper(unsigned long long z) {
  z ^= (z >> 29) & 0x5555555555555555ull;
  z ^= (z << 17) & 0x71D67FFFEDA60000ull;
  z ^= (z << 37) & 0xFFF7EEE000000000ull;
  z ^= (z >> 43);
  return z;
}
// x[0..312) -> next block of the recurrence, in place, by all threads of the CTA
__device__ __forceinline__ void mt64_twist_block(unsigned long long *x) {
  constexpr unsigned long long UPPER = 0xFFFFFFFF80000000ull, LOWER = 0x7FFFFFFFull;
  constexpr unsigned long long A = 0xB5026F5AA96619E9ull;
  auto mix = [&](unsigned long long a, unsigned long long b) {
    const unsigned long long y = (a & UPPER) | (b & LOWER);
    return (y >> 1) ^ ((y & 1ull) ? A : 0ull);
  };
  unsigned long long v[2];
  // k in [0, 156): x[k] = x[k + 156] ^ mix(x[k], x[k + 1]), all operands old
  for (int r = 0, k = threadIdx.x; r < 2; ++r, k += kSerialThreads)
    if (k < 156) v[r] = x[k + 156] ^ mix(x[k], x[k + 1]);
  __syncthreads();
  for (int r = 0, k = threadIdx.x; r < 2; ++r, k += kSerialThreads)
    if (k < 156) x[k] = v[r];
  __syncthreads();
  // k in [156, 311): x[k] = x[k - 156] (new) ^ mix(x[k], x[k + 1]) (old)
  for (int r = 0, k = 156 + threadIdx.x; r < 2; ++r, k += kSerialThreads)
    if (k < 311) v[r] = x[k - 156] ^ mix(x[k], x[k + 1]);
  __syncthreads();
  for (int r = 0, k = 156 + threadIdx.x; r < 2; ++r, k += kSerialThreads)
    if (k < 311) x[k] = v[r];
  if (threadIdx.x == 0) x[311] = x[155] ^ mix(x[311], x[0]);  // x[311] still old, x[0] and x[155] new
  __syncthreads();
}

__global__ void __launch_bounds__(kSerialThreads) k_serial_reference(SerialArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_done;
  const int chain = blockIdx.x;
  unsigned long long *mt = reinterpret_cast<unsigned long long *>(smem_raw);      // recurrence state
  unsigned long long *out = mt + 312;                                              // tempered words
  double *tab_s = reinterpret_cast<double *>(out + 312);                           // dE[16], prob[16]
  uint8_t *lat_s = reinterpret_cast<uint8_t *>(tab_s + 32);
  const NaturalShape s = A.shape;
  uint8_t *nat_g = A.nat + (long long)chain * s.n_sites;
  MT64State *eng = A.engines + chain;

  for (int i = threadIdx.x; i < 312; i += blockDim.x) {
    mt[i] = eng->x[i];
    out[i] = mt64_temper(mt[i]);
  }
  if (threadIdx.x < 16) {
    tab_s[threadIdx.x] = A.tabs[chain].dE[threadIdx.x];
    tab_s[16 + threadIdx.x] = A.tabs[chain].prob[threadIdx.x];
  }
  if (A.use_smem)
    for (long long i = threadIdx.x; i < s.n_sites; i += blockDim.x) lat_s[i] = nat_g[i];
  if (threadIdx.x == 0) s_done = (A.n_passes <= 0);
  __syncthreads();

  // state of thread 0's walk (kept in registers across the pauses)
  uint8_t *nat = A.use_smem ? lat_s : nat_g;
  int pos = eng->pos;
  long long ones = 0, B = 0, pass = 0, step = 0, slot = 0, l = 0;
  unsigned long long n_acc = 0;
  int n_up = 0, b = 0, idx = 0;
  bool need_real = false;
  const int z = 2 * s.dim;
  const bool small = s.n_sites < (1ll << 31);
  const unsigned long long range = (unsigned long long)s.n_sites;
  const unsigned long long lemire_thr = (0ull - range) % range;
  if (threadIdx.x == 0) {
    ones = A.cur_sb[2 * chain];
    B = A.cur_sb[2 * chain + 1];
  }

  while (!s_done) {
    if (pos >= 312) {  // block-uniform: pos is only advanced by thread 0 and published below
      mt64_twist_block(mt);
      for (int i = threadIdx.x; i < 312; i += blockDim.x) out[i] = mt64_temper(mt[i]);
      pos = 0;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      bool done = false;
      while (pos < 312) {
        const unsigned long long u = out[pos++];
        bool accept;
        if (!need_real) {
          // libstdc++ 13 bits/uniform_int_dist.h:252-281 (Lemire, 128-bit product)
          const unsigned long long low = u * range;
          if (low < range && low < lemire_thr) continue;  // redraw
          l = (long long)__umul64hi(u, range);
          n_up = small ? natural_n_up32(nat, s.n0, s.n1, s.n2, s.dim, (unsigned int)l)
                       : natural_n_up(nat, s, l);
          b = nat[l];
          idx = 2 * n_up + b;
          accept = tab_s[idx] < 0.0;
          if (!accept) {
            need_real = true;  // metropolis.hh:28-34: the uniform is drawn only when dE >= 0
            continue;
          }
        } else {
          // bits/random.tcc:3349-3381 (one 64-bit word for 53 bits), then r*(b-a)+a
          double r = __dmul_rn(__ull2double_rn(u), 5.42101086242752217003726400434970855712890625e-20);
          if (r >= 1.0) r = 0.99999999999999988897769753748434595763683319091796875;
          r = __dadd_rn(__dmul_rn(r, __dsub_rn(1.0, 0.0)), 0.0);
          accept = r < tab_s[16 + idx];
          need_real = false;
        }
        if (accept) {
          nat[l] = (uint8_t)(b ^ 1);
          ++n_acc;
          const int ds = -2 * (2 * b - 1);  // new - old
          ones += b ? -1 : 1;
          B += (long long)ds * (2 * n_up - z);
        }
        if (++step == s.n_sites) {  // basic_occupation_metropolis.hh:399-411
          step = 0;
          if (A.sample_period > 0 && A.series && ((A.pass_base + pass + 1) % A.sample_period) == 0) {
            long long *dst = A.series + (long long)chain * A.series_chain_stride + 2 * slot;
            dst[0] = ones;
            dst[1] = B;
            ++slot;
          }
          if (++pass == A.n_passes) {
            done = true;
            break;
          }
        }
      }
      if (done) s_done = 1;
    }
    // every thread learns whether the block of words was used up (pos == 312) or the run ended
    __syncthreads();
    if (!s_done) pos = 312;
  }
  if (threadIdx.x == 0) {
    eng->pos = pos;
    A.cur_sb[2 * chain] = ones;
    A.cur_sb[2 * chain + 1] = B;
    A.n_accept[chain] += n_acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 312; i += blockDim.x) eng->x[i] = mt[i];
  if (A.use_smem)
    for (long long i = threadIdx.x; i < s.n_sites; i += blockDim.x) nat_g[i] = lat_s[i];
}

// draws through the device engine, for the RNG parity probe
__global__ void k_rng_draw(MT64State *eng, int n, const long long *int_max,
                           const double *real_max, const uint8_t *is_real,
                           long long *int_out, double *real_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int pos = eng->pos;
  for (int i = 0; i < n; ++i) {
    if (is_real[i]) {
      real_out[i] = mt64_uniform_real(eng->x, pos, real_max[i]);
      int_out[i] = 0;
    } else {
      const unsigned long long mx = (unsigned long long)int_max[i];
      // range == 2^64 is the identity mapping (uniform_int_dist.h:283-286)
      int_out[i] = (mx == ~0ull) ? (long long)mt64_next(eng->x, pos)
                                 : (long long)mt64_uniform_int(eng->x, pos, mx + 1ull);
      real_out[i] = 0.0;
    }
  }
  eng->pos = pos;
}

// ---------------------------------------------------------------------------
// Sample series -> intensive doubles, with the reference's expression order
// (model.hh:266-270, :412-422, :293-295; basic_semigrand_canonical.hh:165-174).
// Explicit round-to-nearest intrinsics keep the compiler from contracting
// mul+sub into an FMA, which would change the last bit.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void intensive_from_sums(long long ones, long long B,
                                                    long long N, double J, double mu,
                                                    double &x, double &ef, double &ep) {
  const long long S = 2 * ones - N;
  const double e_formation = __dmul_rn((double)B, -J);
  const double Nx = __ddiv_rn((double)(N + S), 2.0);
  x = __ddiv_rn(Nx, (double)N);
  ef = __ddiv_rn(e_formation, (double)N);
  ep = __ddiv_rn(__dsub_rn(e_formation, __dmul_rn(mu, Nx)), (double)N);
}

// series (ones,B) int64 pairs -> three double columns, for samples [first, first+count)
__global__ void k_series_to_doubles(const long long *__restrict__ sb, long long first,
                                    long long count, long long N, const ChainTables *tab,
                                    double *x, double *ef, double *ep) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double a, b, c;
  intensive_from_sums(sb[2 * (first + i)], sb[2 * (first + i) + 1], N, tab->J, tab->mu, a, b, c);
  x[first + i] = a;
  ef[first + i] = b;
  ep[first + i] = c;
}

// ---------------------------------------------------------------------------
// The use_nlist = false energy form on the device (model.hh:273-285):
//   E = sum_i (-J * dot(row i, row i+1)) + sum_j (-J * dot(col j, col j+1)),
// every term rounded and accumulated in double, rows first.  The dots are
// integers: k_line_xor_planes counts, straight from the colour planes, the
// unequal neighbour pairs of every "row" (fixed i) and "column" (fixed j) at a
// sampled pass; dot = length - 2 * count.  k_nonlist_to_doubles then walks the
// reference's accumulation order, one thread per sample.
//   even row i = 2p   : planes 0 and 1 at (p, j)
//   odd row i = 2p+1  : plane 1-(j&1) at (p, j) with plane (j&1) at (p+1, j)
//   column j          : plane c at (p, j) with plane 1-c at (p, j+1), both c
// A thread owns 16 bytes of p and walks a strip of at most 255 columns with
// per-byte counters packed in registers.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_line_xor_planes(LatticeView L, int chain, int js,
                                                         int *__restrict__ row_cnt,
                                                         int *__restrict__ col_cnt) {
  const int h = L.h, n1 = L.n1;
  const int V = h >> 4;
  const int n_strips = (n1 + js - 1) / js;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < (long long)V * n_strips;
  const int v = active ? (int)(t % V) : 0;
  const int strip = active ? (int)(t / V) : 0;
  const int p0 = v << 4;
  const int jbeg = strip * js;
  const int jend = active ? min(jbeg + js, n1) : jbeg;
  const uint8_t *P0 = L.planes + (long long)chain * L.chain_stride;
  const uint8_t *P1 = P0 + L.plane_stride;
  const bool warp_uniform = (V & 31) == 0;  // a warp then lies inside one strip
  const int p_above = (p0 + 16 == h) ? 0 : p0 + 16;
  uint4 ev = make_uint4(0u, 0u, 0u, 0u), od = ev;
  uint4 a0 = ev, a1 = ev;
  if (jbeg < jend) {
    a0 = ld16(P0 + (long long)h * jbeg + p0);
    a1 = ld16(P1 + (long long)h * jbeg + p0);
  }
  for (int j = jbeg; j < jend; ++j) {
    const int jn = (j + 1 == n1) ? 0 : j + 1;
    const uint4 b0 = ld16(P0 + (long long)h * jn + p0);
    const uint4 b1 = ld16(P1 + (long long)h * jn + p0);
    ev.x += a0.x ^ a1.x;
    ev.y += a0.y ^ a1.y;
    ev.z += a0.z ^ a1.z;
    ev.w += a0.w ^ a1.w;
    const uint4 A = (j & 1) ? a0 : a1;  // plane 1-(j&1) at p
    const uint4 B = (j & 1) ? a1 : a0;  // plane (j&1), taken at p+1
    const uint32_t eb = ((j & 1) ? P1 : P0)[(long long)h * j + p_above];
    const uint4 Bs = shift_down_1(B, eb);
    od.x += A.x ^ Bs.x;
    od.y += A.y ^ Bs.y;
    od.z += A.z ^ Bs.z;
    od.w += A.w ^ Bs.w;
    int cnt = bytesum(a0.x ^ b1.x) + bytesum(a0.y ^ b1.y) + bytesum(a0.z ^ b1.z) +
              bytesum(a0.w ^ b1.w) + bytesum(a1.x ^ b0.x) + bytesum(a1.y ^ b0.y) +
              bytesum(a1.z ^ b0.z) + bytesum(a1.w ^ b0.w);
    if (warp_uniform) {
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if ((threadIdx.x & 31) == 0) atomicAdd(&col_cnt[j], cnt);
    } else {
      atomicAdd(&col_cnt[j], cnt);
    }
    a0 = b0;
    a1 = b1;
  }
  if (jbeg < jend) {
    const uint32_t e[4] = {ev.x, ev.y, ev.z, ev.w}, o[4] = {od.x, od.y, od.z, od.w};
#pragma unroll
    for (int b = 0; b < 16; ++b) {
      atomicAdd(&row_cnt[2 * (p0 + b)], (int)((e[b >> 2] >> (8 * (b & 3))) & 0xffu));
      atomicAdd(&row_cnt[2 * (p0 + b) + 1], (int)((o[b >> 2] >> (8 * (b & 3))) & 0xffu));
    }
  }
}

// lines: per sample n0 row counts then n1 column counts.  Overwrites the
// formation / potential energy columns of samples [first, first + count).
__global__ void k_nonlist_to_doubles(const int *__restrict__ lines, long long line_stride,
                                     int n0, int n1, const long long *__restrict__ sb,
                                     long long first, long long count, long long N,
                                     const ChainTables *tab, double *ef, double *ep) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int *ln = lines + (first + i) * line_stride;
  const double mJ = -tab->J;
  double e = 0.0;
  for (int r = 0; r < n0; ++r) e = __dadd_rn(e, __dmul_rn(mJ, (double)(n1 - 2 * ln[r])));
  for (int c = 0; c < n1; ++c) e = __dadd_rn(e, __dmul_rn(mJ, (double)(n0 - 2 * ln[n0 + c])));
  const long long ones = sb[2 * (first + i)];
  const long long S = 2 * ones - N;
  const double Nx = __ddiv_rn((double)(N + S), 2.0);
  ef[first + i] = __ddiv_rn(e, (double)N);
  ep[first + i] = __ddiv_rn(__dsub_rn(e, __dmul_rn(tab->mu, Nx)), (double)N);
}

// ---------------------------------------------------------------------------
// Series statistics (src/casm/monte/BasicStatistics.cc:24-48, :114-131;
// include/casm/monte/misc/math.hh:21-39).  One CTA per series.  mean and
// variance by block reduction; the lag search evaluates lag-k autocovariances
// in order until |cov_k / cov_0| <= 0.5.
// out: {mean, variance, f_autocorr, precision}, k_star
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256(double v, double *sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += sh[w];
  return t;
}

struct SeriesJob {
  const double *x;  // series start (already offset to `first`)
  long long n;
};

__global__ void __launch_bounds__(256) k_series_stats(const SeriesJob *jobs, double z_conf,
                                                      double *out4, long long *k_star) {
  __shared__ double sh[8];
  const SeriesJob job = jobs[blockIdx.x];
  const double *x = job.x;
  const long long N = job.n;
  double *out = out4 + 4 * blockIdx.x;
  if (N <= 0) {
    if (threadIdx.x == 0) {
      out[0] = out[1] = out[2] = out[3] = 0.0;
      k_star[blockIdx.x] = -2;
    }
    return;
  }
  double s = 0.0;
  for (long long i = threadIdx.x; i < N; i += blockDim.x) s += x[i];
  const double mean = block_sum_256(s, sh) / (double)N;
  s = 0.0;
  for (long long i = threadIdx.x; i < N; i += blockDim.x) {
    const double d = x[i] - mean;
    s += d * d;
  }
  const double var = block_sum_256(s, sh) / (double)N;

  double f = 1.0;
  long long ks = 0;
  if (!(fabs(var / mean) < 1e-8 || var == 0.0)) {
    f = 1.7976931348623157e308;
    ks = -1;
    for (long long k = 1; k < N; ++k) {
      const long long m = N - k;
      double c = 0.0;
      for (long long i = threadIdx.x; i < m; i += blockDim.x)
        c += (x[i] - mean) * (x[i + k] - mean);
      const double cov = block_sum_256(c, sh) / (double)m;
      if (fabs(cov / var) <= 0.5) {
        const double rho = pow(2.0, -1.0 / (double)k);
        f = (1.0 + rho) / (1.0 - rho);
        ks = k;
        break;
      }
    }
  }
  if (threadIdx.x == 0) {
    out[0] = mean;
    out[1] = var;
    out[2] = f;
    out[3] = z_conf * sqrt(f * var / (double)N);
    k_star[blockIdx.x] = ks;
  }
}

// ---------------------------------------------------------------------------
// Weighted observations (src/casm/monte/BasicStatistics.cc:50-73, :144-188;
// include/casm/monte/misc/math.hh:46-55): the series is a time series of
// unequal intervals.  One CTA per job.
//  * W = sum(w) and the two-pointer walk of `resample` are sequential
//    floating-point recurrences whose comparisons select samples, so thread 0
//    evaluates them in the reference order (explicit _rn operations);
//  * the weighted mean / variance and the lag search on the resampled series
//    are block reductions like k_series_stats.
// method 1: mean and variance from the weighted samples, only the
// autocorrelation factor from the resampled series (rho = 2^(-1/(k*increment)));
// method 2: everything from the resampled series.
// out: {mean, variance, f_autocorr, precision, W}, k_star
// ---------------------------------------------------------------------------
struct WeightedJob {
  const double *x;
  const double *w;
  long long n;
  double *resampled;  // n_resamples doubles of scratch / output
  long long n_resamples;
  int method;         // 1 or 2; 0 = resample only (W taken from weight_sum)
  double weight_sum;  // used when method == 0
};

__device__ __forceinline__ void block_lag_search(const double *x, long long N, double mean,
                                                 double var, double increment, double *sh,
                                                 double *f_out, long long *ks_out) {
  double f = 1.0;
  long long ks = 0;
  if (!(fabs(var / mean) < 1e-8 || var == 0.0)) {
    f = 1.7976931348623157e308;
    ks = -1;
    for (long long k = 1; k < N; ++k) {
      const long long m = N - k;
      double c = 0.0;
      for (long long i = threadIdx.x; i < m; i += blockDim.x)
        c += (x[i] - mean) * (x[i + k] - mean);
      const double cov = block_sum_256(c, sh) / (double)m;
      if (fabs(cov / var) <= 0.5) {
        const double rho = pow(2.0, __ddiv_rn(-1.0, __dmul_rn((double)k, increment)));
        f = (1.0 + rho) / (1.0 - rho);
        ks = k;
        break;
      }
    }
  }
  *f_out = f;
  *ks_out = ks;
}

__global__ void __launch_bounds__(256) k_series_stats_weighted(const WeightedJob *jobs,
                                                               double z_conf, double *out5,
                                                               long long *k_star) {
  __shared__ double sh[8];
  __shared__ double sW;
  const WeightedJob job = jobs[blockIdx.x];
  const double *x = job.x, *w = job.w;
  const long long n = job.n, R = job.n_resamples;
  double *eq = job.resampled;
  double *out = out5 + 5 * blockIdx.x;
  if (threadIdx.x == 0) {
    double W = job.weight_sum;
    if (job.method != 0) {
      W = 0.0;
      for (long long i = 0; i < n; ++i) W = __dadd_rn(W, w[i]);
    }
    sW = W;
    // resample (BasicStatistics.cc:57-72); j is kept inside the series
    const double increment = __ddiv_rn(W, (double)R);
    long long j = 0;
    double W_j = 0.0;
    for (long long i = 0; i < R; ++i) {
      const double W_target = __dmul_rn((double)i, increment);
      while (j < n - 1 && __dadd_rn(W_j, w[j]) < W_target) {
        W_j = __dadd_rn(W_j, w[j]);
        ++j;
      }
      eq[i] = x[j];
    }
  }
  __syncthreads();
  if (job.method == 0) return;
  const double W = sW;
  const double increment = __ddiv_rn(W, (double)R);

  // statistics of the resampled series
  double s = 0.0;
  for (long long i = threadIdx.x; i < R; i += blockDim.x) s += eq[i];
  const double mean_eq = block_sum_256(s, sh) / (double)R;
  s = 0.0;
  for (long long i = threadIdx.x; i < R; i += blockDim.x) {
    const double d = eq[i] - mean_eq;
    s += d * d;
  }
  const double var_eq = block_sum_256(s, sh) / (double)R;
  double f;
  long long ks;
  block_lag_search(eq, R, mean_eq, var_eq, job.method == 1 ? increment : 1.0, sh, &f, &ks);

  double mean = mean_eq, var = var_eq, denom = (double)R;
  if (job.method == 1) {
    s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += x[i] * w[i];
    mean = block_sum_256(s, sh) / W;
    s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const double d = x[i] - mean;
      s += w[i] * d * d;
    }
    var = block_sum_256(s, sh) / W;
    denom = W;
  }
  if (threadIdx.x == 0) {
    out[0] = mean;
    out[1] = var;
    out[2] = f;
    out[3] = z_conf * sqrt(f * var / denom);
    out[4] = W;
    k_star[blockIdx.x] = ks;
  }
}

// weighted_observation(i) = x(i) * ((N / W) * w(i)), W summed in order
// (src/casm/monte/checks/EquilibrationCheck.cc:145-158); two launches:
// k_weight_factor (one thread) then k_apply_weight_factor.
__global__ void k_weight_factor(const double *__restrict__ w, long long n, double *factor) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double W = 0.0;
  for (long long i = 0; i < n; ++i) W = __dadd_rn(W, w[i]);
  *factor = __ddiv_rn((double)n, W);
}
__global__ void k_apply_weight_factor(const double *__restrict__ x, const double *__restrict__ w,
                                      long long n, const double *__restrict__ factor,
                                      double *__restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = __dmul_rn(x[i], __dmul_rn(*factor, w[i]));
}

// Equilibration check (src/casm/monte/checks/EquilibrationCheck.cc:50-117).  One CTA
// per series.  "All samples equal" is decided in parallel and the two initial
// partition sums are block reductions (the reference takes them with Eigen's
// .sum(), whose order is not a sequential one either).  The scan that follows is a
// floating-point recurrence on the running sums (:93-103) and keeps the reference's
// order: it is walked in chunks; all threads first bring the two runs of samples the
// chunk consumes into shared memory, thread 0 walks the recurrence recording
// (sum1, sum2) before every step, and all threads then evaluate the loop condition
// of the recorded steps -- two double divisions each, which dominated when one
// thread did everything (1.7 ms per check of 10^4 samples that never equilibrate,
// 0.44 ms now) -- and find the step at which the reference's loop stops.  Shared
// memory is a few KiB whatever the series length, so the kernel runs next to a
// sweep kernel (see cmg_mark).
constexpr int kEquilThreads = 256;
constexpr int kEquilChunk = 512;
__global__ void __launch_bounds__(kEquilThreads) k_series_equilibration(const SeriesJob *jobs,
                                                                        int n_jobs, double prec,
                                                                        int *is_eq, long long *n_eq) {
  __shared__ int s_differs;
  __shared__ int s_stop;  // first step of the chunk at which the loop condition is false
  __shared__ double s_sum1[kEquilChunk], s_sum2[kEquilChunk], s_red[8];
  __shared__ double s_x1[kEquilChunk], s_x2[kEquilChunk / 2 + 1];  // x[start1 + k], x[start2 + m]
  __shared__ double s_end1, s_end2;
  const int jb = blockIdx.x;
  if (jb >= n_jobs) return;
  const double *x = jobs[jb].x;
  const long long N = jobs[jb].n;
  if (N <= 0) {
    if (threadIdx.x == 0) {
      is_eq[jb] = 0;
      n_eq[jb] = 0;
    }
    return;
  }
  if (threadIdx.x == 0) s_differs = 0;
  __syncthreads();
  const double x0 = x[0];
  const double eps = (x0 == 0.0) ? 1e-8 : fabs(x0) * 1e-8;
  const bool even0 = ((N % 2) == 0);
  const long long start2_0 = even0 ? N / 2 : (N / 2) + 1;
  bool differs = false;
  double p1 = 0.0, p2 = 0.0;
  for (long long i = threadIdx.x; i < N; i += kEquilThreads) {
    const double v = x[i];
    differs |= fabs(v - x0) > eps;
    if (i < start2_0)
      p1 += v;
    else
      p2 += v;
  }
  if (differs) s_differs = 1;
  double sum1 = block_sum_256(p1, s_red);
  double sum2 = block_sum_256(p2, s_red);
  if (!s_differs) {  // all samples (approximately) equal
    if (threadIdx.x == 0) {
      is_eq[jb] = 1;
      n_eq[jb] = 0;
    }
    return;
  }

  // state before the first step of the chunk (kept by every thread; thread 0 advances the sums)
  long long start1 = 0, start2 = start2_0;
  const bool is_even = even0;  // parity of the chunk's first step (chunks hold an even number of steps)
  for (;;) {
    // the samples this chunk can consume: x[start1 .. start1 + chunk) and x[start2 .. start2 + chunk/2]
    for (int k = threadIdx.x; k < kEquilChunk; k += kEquilThreads) s_x1[k] = (start1 + k < N) ? x[start1 + k] : 0.0;
    for (int k = threadIdx.x; k < kEquilChunk / 2 + 1; k += kEquilThreads) s_x2[k] = (start2 + k < N) ? x[start2 + k] : 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
      s_stop = kEquilChunk;
      double a = sum1, b = sum2;
      // EquilibrationCheck.cc:93-103, two steps (one of each parity) per trip: the parity
      // pattern of a chunk is fixed, so the trip has no branch and its loads do not depend on
      // the sums -- what is left on the critical path is three dependent additions per trip
      const long long left = N - 2 - start1;
      const int steps = (int)(left < (long long)kEquilChunk ? (left > 0 ? left : 0) : kEquilChunk);
      int k = 0, m = 0;
      if (is_even) {
#pragma unroll 4
        for (; k + 2 <= steps; k += 2, ++m) {
          const double xa = s_x1[k], xb = s_x1[k + 1], x2 = s_x2[m];
          s_sum1[k] = a;
          s_sum2[k] = b;
          a = __dadd_rn(__dsub_rn(a, xa), x2);
          b = __dsub_rn(b, x2);
          s_sum1[k + 1] = a;
          s_sum2[k + 1] = b;
          a = __dsub_rn(a, xb);
        }
        if (k < steps) {
          s_sum1[k] = a;
          s_sum2[k] = b;
          a = __dadd_rn(__dsub_rn(a, s_x1[k]), s_x2[m]);
          b = __dsub_rn(b, s_x2[m]);
        }
      } else {
#pragma unroll 4
        for (; k + 2 <= steps; k += 2, ++m) {
          const double xa = s_x1[k], xb = s_x1[k + 1], x2 = s_x2[m];
          s_sum1[k] = a;
          s_sum2[k] = b;
          a = __dsub_rn(a, xa);
          s_sum1[k + 1] = a;
          s_sum2[k + 1] = b;
          a = __dadd_rn(__dsub_rn(a, xb), x2);
          b = __dsub_rn(b, x2);
        }
        if (k < steps) {
          s_sum1[k] = a;
          s_sum2[k] = b;
          a = __dsub_rn(a, s_x1[k]);
        }
      }
      s_end1 = a;
      s_end2 = b;
    }
    __syncthreads();
    // the loop condition (:91-92) of the chunk's steps, in parallel
    int first = kEquilChunk;
    for (int k = threadIdx.x; k < kEquilChunk; k += kEquilThreads) {
      const long long s1 = start1 + k;
      // start2 has advanced once per even step taken so far
      const long long s2 = start2 + (is_even ? (k + 1) / 2 : k / 2);
      bool go = s1 < N - 2;
      if (go)
        go = fabs(__dsub_rn(__ddiv_rn(s_sum1[k], (double)(s2 - s1)), __ddiv_rn(s_sum2[k], (double)(N - s2)))) > prec;
      if (!go) {
        first = k;
        break;
      }
    }
    if (first < kEquilChunk) atomicMin(&s_stop, first);
    __syncthreads();
    const int stop = s_stop;
    if (stop < kEquilChunk) {
      // the reference's loop ends before step `stop`; its sums are the recorded ones
      // (or, when the scan ran out of samples, the state after the last step taken)
      const bool recorded = start1 + stop < N - 2;
      sum1 = recorded ? s_sum1[stop] : s_end1;
      sum2 = recorded ? s_sum2[stop] : s_end2;
      start2 += is_even ? (stop + 1) / 2 : stop / 2;
      start1 += stop;
      break;
    }
    sum1 = s_end1;
    sum2 = s_end2;
    start2 += kEquilChunk / 2;
    start1 += kEquilChunk;
    __syncthreads();  // the buffers are rewritten by the next chunk
  }
  if (threadIdx.x != 0) return;
  const double mean_tot = __ddiv_rn(__dadd_rn(sum1, sum2), (double)(N - start1));
  if (x[start1] < mean_tot) {
    while (x[start1] < mean_tot && start1 < N - 1) start1++;
  } else {
    while (x[start1] > mean_tot && start1 < N - 1) start1++;
  }
  is_eq[jb] = (start1 < N - 1) ? 1 : 0;
  n_eq[jb] = start1;
}

// One completion check on the device-resident series without a host round trip in
// the middle (CompletionCheck::_check_convergence, include/casm/monte/checks/
// CompletionCheck.hh:353-376): after k_series_equilibration has filled is_eq / n_eq
// for the requested components (in the caller's order), this builds the jobs of
// the statistics pass: if every component up to the first failure equilibrated,
// tail(count - max n_eq) of each series (ConvergenceCheck.hh:139-184); otherwise
// empty jobs.
__global__ void k_make_tail_jobs(const SeriesJob *eq_jobs, int n_jobs, const int *is_eq,
                                 const long long *n_eq, SeriesJob *stat_jobs, long long *n_stats) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  bool all = true;
  long long first = 0;
  for (int i = 0; i < n_jobs; ++i) {
    all = all && is_eq[i] != 0;
    if (n_eq[i] > first) first = n_eq[i];
  }
  const long long count = n_jobs > 0 ? eq_jobs[0].n : 0;
  const bool ok = all && first < count;
  for (int i = 0; i < n_jobs; ++i) {
    stat_jobs[i].x = eq_jobs[i].x + (ok ? first : 0);
    stat_jobs[i].n = ok ? count - first : 0;
  }
  *n_stats = ok ? count - first : 0;
}

// ---------------------------------------------------------------------------
// Supercell index conversions, diagonal transformation matrix
// (src/casm/monte/Conversions.cc:181-229 forwarding to
// xtal::UnitCellCoordIndexConverter): l = b*n_unitcells + i + n0*(j + n1*k)
// ---------------------------------------------------------------------------
__global__ void k_conv_l_to_bijk(long long n0, long long n1, long long n2,
                                 const long long *__restrict__ l, long long count,
                                 long long *bijk) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const long long vol = n0 * n1 * n2;
  const long long v = l[t];
  const long long u = v % vol;
  bijk[4 * t + 0] = v / vol;
  bijk[4 * t + 1] = u % n0;
  bijk[4 * t + 2] = (u / n0) % n1;
  bijk[4 * t + 3] = u / (n0 * n1);
}
__device__ __forceinline__ long long floor_mod(long long a, long long m) {
  long long r = a % m;
  return r < 0 ? r + m : r;
}
__global__ void k_conv_bijk_to_l(long long n0, long long n1, long long n2,
                                 const long long *__restrict__ bijk, long long count,
                                 long long *l) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const long long vol = n0 * n1 * n2;
  l[t] = bijk[4 * t] * vol + floor_mod(bijk[4 * t + 1], n0) +
         n0 * (floor_mod(bijk[4 * t + 2], n1) + n1 * floor_mod(bijk[4 * t + 3], n2));
}

// General integer transformation matrix (include/casm_monte_b200/snf.hh restates
// xtal::UnitCellCoordIndexConverter): unit cell of index ix = U * (ix % s0,
// (ix / s0) % s1, ix / (s0 s1)) brought within the supercell; inverse through
// U^-1 and mod s.  Matrices row-major.
struct ConvGeneral {
  long long T[9], adjT[9], U[9], Uinv[9];
  long long detT, s[3], n_unitcells;
};
__device__ __forceinline__ long long floor_div_ll(long long x, long long y) {
  long long q = x / y, r = x % y;
  return (r != 0 && ((r < 0) != (y < 0))) ? q - 1 : q;
}
__global__ void k_conv_general_l_to_bijk(ConvGeneral P, const long long *__restrict__ l, long long count,
                                         long long *bijk) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const long long v = l[t];
  const long long ix = v % P.n_unitcells;
  const long long mnp[3] = {ix % P.s[0], (ix / P.s[0]) % P.s[1], ix / (P.s[0] * P.s[1])};
  long long ijk[3], f[3];
  for (int r = 0; r < 3; ++r) ijk[r] = P.U[3 * r] * mnp[0] + P.U[3 * r + 1] * mnp[1] + P.U[3 * r + 2] * mnp[2];
  for (int r = 0; r < 3; ++r)
    f[r] = floor_div_ll(P.adjT[3 * r] * ijk[0] + P.adjT[3 * r + 1] * ijk[1] + P.adjT[3 * r + 2] * ijk[2], P.detT);
  bijk[4 * t] = v / P.n_unitcells;
  for (int r = 0; r < 3; ++r)
    bijk[4 * t + 1 + r] = ijk[r] - (P.T[3 * r] * f[0] + P.T[3 * r + 1] * f[1] + P.T[3 * r + 2] * f[2]);
}
__global__ void k_conv_general_bijk_to_l(ConvGeneral P, const long long *__restrict__ bijk, long long count,
                                         long long *l) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const long long *in = bijk + 4 * t;
  long long mnp[3];
  for (int r = 0; r < 3; ++r)
    mnp[r] = floor_mod(P.Uinv[3 * r] * in[1] + P.Uinv[3 * r + 1] * in[2] + P.Uinv[3 * r + 2] * in[3], P.s[r]);
  l[t] = in[0] * P.n_unitcells + mnp[0] + P.s[0] * (mnp[1] + P.s[1] * mnp[2]);
}

}  // namespace cmg

// ===========================================================================
// k-state lattice model served by the general multi-species proposal tables
// (SURVEY 8f rank 3).  K <= 4 species per site, occupation byte = occupation
// index; nearest-neighbour pair energy V[a][b] and exchange potential mu[s]:
//     potential = sum_<ij> V[o_i][o_j] - sum_i mu[o_i].
// The host builds, per chain, the table of dPhi / exp(-dPhi*beta) / 32-bit
// threshold for every (from, to, neighbour configuration); configuration index
// = sum_{s>=1} n_s * (z+1)^(s-1), n_s = number of neighbours holding species s.
// Two update orders, as for the Ising path:
//   checkerboard: every site of a colour proposes one of its K-1 other species
//     (one Philox call per site: word 0 chooses the species, word 1 is the
//     acceptance uniform);
//   serial reference: the reference's general proposal machinery --
//     OccCandidateList / OccLocation / propose_semigrand_canonical_event
//     (include/casm/monte/events/OccEventProposal.hh:260-348,
//     src/casm/monte/events/OccLocation.cc:39-116, :253-283) -- walked by one
//     thread per chain on the reference's mt19937_64 stream, trajectory-exact.
// ===========================================================================
namespace cmg {

constexpr int kMaxSpecies = 4;
constexpr int kMaxKCfg = 343;  // (2*3+1)^(4-1)
constexpr int kMaxKEntries = kMaxSpecies * kMaxSpecies * kMaxKCfg;

struct KStateTables {
  double dPhi[kMaxKEntries];
  double prob[kMaxKEntries];
  uint32_t thr_m1[kMaxKEntries];
  uint8_t never[kMaxKEntries];
  int K, z, n_cfg, valid;
};
__device__ __forceinline__ int kstate_entry(const KStateTables *t, int from, int to, int cfg) {
  return (from * t->K + to) * t->n_cfg + cfg;
}
// (z+1)^(s-1) for s = 0..3 (weight of species 0 is 0: its count is implied)
__device__ __forceinline__ int kstate_weight(int s, int z) {
  return s == 0 ? 0 : s == 1 ? 1 : s == 2 ? (z + 1) : (z + 1) * (z + 1);
}

struct KSweepArgs {
  LatticeView L;
  const KStateTables *tabs;  // [chain]
  unsigned long long *n_accept;
  unsigned long long pass;
  uint32_t rk[20];
  int colour;
  int chain_offset;
};

// Coloured half-sweep.  One Philox call serves the two sites p = 2g, 2g + 1 of a column
// (group G = g + ceil(h/2) * (j + n1 * k); words 0, 1 for the even site, 2, 3 for the odd one:
// species choice and acceptance uniform).  Block (32, 8): a warp takes 32 consecutive groups
// of one column, the block eight columns per trip and kKStateTrips trips (no division or
// modulo per site); grid.z = k + n2 * chain.  The chain's threshold table is staged in shared
// memory when it fits (every model but K = 4 in 3-d), accepted counts are reduced per block:
// one atomic per block (the one-atomic-per-warp form was bound by same-address atomics in
// L2).  Launched as a programmatic dependent of the half-sweep before it (pdl_wait).
constexpr int kKStateTrips = 8;
constexpr int kKStateShared = 2048;
__global__ void __launch_bounds__(256, 3) k_kstate_halfsweep(KSweepArgs A) {
  __shared__ uint32_t s_thr[kKStateShared];
  __shared__ uint8_t s_never[kKStateShared];
  __shared__ unsigned int s_acc;
  const LatticeView &L = A.L;
  const int chain = (int)(blockIdx.z / (unsigned)L.n2);
  const int k = (int)(blockIdx.z - (unsigned)chain * (unsigned)L.n2);
  const int g = (int)(blockIdx.x * 32u + threadIdx.x);
  const int hh = (L.h + 1) >> 1;
  const KStateTables *tab = A.tabs + chain;
  const int K = tab->K, n_cfg = tab->n_cfg;
  const int n_entries = K * K * n_cfg;
  const bool staged = n_entries <= kKStateShared;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (staged)
    for (int e = tid; e < n_entries; e += 256) {
      s_thr[e] = tab->thr_m1[e];
      s_never[e] = tab->never[e];
    }
  if (tid == 0) s_acc = 0;
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();
  uint8_t *C = L.planes + (long long)chain * L.chain_stride + (long long)A.colour * L.plane_stride;
  const uint8_t *O = L.planes + (long long)chain * L.chain_stride + (long long)(1 - A.colour) * L.plane_stride;
  const int z = 2 * L.dim;
  const uint32_t chain_word = (uint32_t)(chain + A.chain_offset) << 8;
  const long long rowk = (long long)L.n1 * k;
  const int km = (k == 0) ? L.n2 - 1 : k - 1, kp = (k == L.n2 - 1) ? 0 : k + 1;
  unsigned int acc = 0;
  // weights of the four species, one per byte: branch-free table in a register (the ternary
  // chain compiled to two divergent branches per neighbour byte, which serialised the loads)
  const uint32_t wpack = (1u << 8) | ((uint32_t)(z + 1) << 16) | ((uint32_t)((z + 1) * (z + 1)) << 24);
  auto weight = [&](uint32_t sp) { return (wpack >> (8u * sp)) & 0xffu; };
  if (g < hh) {
    const int p0 = 2 * g;
    const bool two = p0 + 1 < L.h;
    const int p1 = two ? p0 + 1 : p0;
    const int pb = (p0 == 0) ? L.h - 1 : p0 - 1;        // below the even site
    const int pa = (p1 == L.h - 1) ? 0 : p1 + 1;        // above the odd site
#pragma unroll 2
    for (int t = 0; t < kKStateTrips; ++t) {
      const int j = (int)((blockIdx.y * kKStateTrips + t) * 8u + threadIdx.y);
      if (j >= L.n1) break;
      const int par = (j + k + A.colour) & 1;  // i = 2p + par
      const int jm = (j == 0) ? L.n1 - 1 : j - 1, jp = (j == L.n1 - 1) ? 0 : j + 1;
      const long long col = (long long)L.h * (j + rowk);
      const uint8_t *__restrict__ Oc = O + col;
      const uint8_t *__restrict__ Om = O + (long long)L.h * (jm + rowk);
      const uint8_t *__restrict__ Op = O + (long long)L.h * (jp + rowk);
      uint8_t *__restrict__ Cc = C + col;
      // every load of the two sites first
      const uint32_t m0 = Om[p0], m1 = Om[p1], q0 = Op[p0], q1 = Op[p1], c0 = Oc[p0], c1 = Oc[p1];
      const uint32_t e0 = Oc[par ? p1 : pb];  // the other i-neighbour of the even site (p1 == p0 + 1 when two)
      const uint32_t e1 = Oc[par ? pa : p0];  // ... of the odd site
      uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 0;
      if (L.dim == 3) {
        const uint8_t *__restrict__ Oa = O + (long long)L.h * (j + (long long)L.n1 * km);
        const uint8_t *__restrict__ Ob = O + (long long)L.h * (j + (long long)L.n1 * kp);
        a0 = Oa[p0], a1 = Oa[p1], b0 = Ob[p0], b1 = Ob[p1];
      }
      const int from0 = Cc[p0], from1 = Cc[p1];
      const uint4 w = site_group_random((unsigned long long)g + (unsigned long long)hh * (unsigned long long)(j + rowk),
                                        chain_word, A.pass, A.colour, 0, A.rk);
      // even site; when the column has an odd number of plane indices its last group holds
      // one site only, whose "other i-neighbour" for par = 1 is p0 + 1 wrapped to 0
      const uint32_t e0x = (!two && par) ? (uint32_t)Oc[(p0 == L.h - 1) ? 0 : p0 + 1] : e0;
      int cfg0 = (int)(weight(m0) + weight(q0) + weight(c0) + weight(e0x));
      int cfg1 = (int)(weight(m1) + weight(q1) + weight(c1) + weight(e1));
      if (L.dim == 3) {
        cfg0 += (int)(weight(a0) + weight(b0));
        cfg1 += (int)(weight(a1) + weight(b1));
      }
      const int jj0 = (int)__umulhi(w.x, (uint32_t)(K - 1)), jj1 = (int)__umulhi(w.z, (uint32_t)(K - 1));
      const int to0 = jj0 + (jj0 >= from0 ? 1 : 0), to1 = jj1 + (jj1 >= from1 ? 1 : 0);
      const int i0 = (from0 * K + to0) * n_cfg + cfg0, i1 = (from1 * K + to1) * n_cfg + cfg1;
      uint32_t thr0, thr1;
      bool nv0, nv1;
      if (staged) {
        thr0 = s_thr[i0], thr1 = s_thr[i1];
        nv0 = s_never[i0], nv1 = s_never[i1];
      } else {
        thr0 = tab->thr_m1[i0], thr1 = tab->thr_m1[i1];
        nv0 = tab->never[i0], nv1 = tab->never[i1];
      }
      if (!nv0 && w.y <= thr0) {
        Cc[p0] = (uint8_t)to0;
        ++acc;
      }
      if (two && !nv1 && w.w <= thr1) {
        Cc[p1] = (uint8_t)to1;
        ++acc;
      }
    }
  }
  acc = __reduce_add_sync(0xffffffffu, acc);
  if (threadIdx.x == 0 && acc) atomicAdd(&s_acc, acc);
  __syncthreads();
  if (tid == 0 && s_acc) atomicAdd(A.n_accept + chain, (unsigned long long)s_acc);
}

// integer observables of a k-state lattice in the natural layout: count[s] and the
// bond-type histogram bonds[a][b], a <= b, over the (+i, +j [, +k]) bonds of every site.
// out: K counts then K*K bond counts (row-major, only a <= b used)
__global__ void __launch_bounds__(256) k_kstate_observables(const uint8_t *__restrict__ nat, NaturalShape s, int K,
                                                            long long *out) {
  __shared__ unsigned long long sh[kMaxSpecies + kMaxSpecies * kMaxSpecies];
  for (int i = threadIdx.x; i < K + K * K; i += blockDim.x) sh[i] = 0ull;
  __syncthreads();
  for (long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x; l < s.n_sites; l += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(l % s.n0);
    const long long r = l / s.n0;
    const int j = (int)(r % s.n1);
    const int k = (int)(r / s.n1);
    const int a = nat[l];
    atomicAdd(&sh[a], 1ull);
    const int ip = (i + 1 == s.n0) ? 0 : i + 1;
    const int jp = (j + 1 == s.n1) ? 0 : j + 1;
    int b = nat[ip + (long long)s.n0 * (j + (long long)s.n1 * k)];
    atomicAdd(&sh[K + min(a, b) * K + max(a, b)], 1ull);
    b = nat[i + (long long)s.n0 * (jp + (long long)s.n1 * k)];
    atomicAdd(&sh[K + min(a, b) * K + max(a, b)], 1ull);
    if (s.dim == 3) {
      const int kp = (k + 1 == s.n2) ? 0 : k + 1;
      b = nat[i + (long long)s.n0 * (j + (long long)s.n1 * kp)];
      atomicAdd(&sh[K + min(a, b) * K + max(a, b)], 1ull);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K + K * K; i += blockDim.x)
    if (sh[i]) atomicAdd(reinterpret_cast<unsigned long long *>(out + i), sh[i]);
}

// values must be occupation indices 0..K-1
__global__ void k_kstate_i32_to_sites(const int32_t *__restrict__ src, uint8_t *base, long long plane_stride,
                                      NaturalShape s, int planar, int K, int *bad) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  const int v = src[l];
  if (v < 0 || v >= K) atomicExch(bad, 1);
  base[site_addr_of(s, l, planar, plane_stride)] = (uint8_t)v;
}
__global__ void k_kstate_sites_to_i32(const uint8_t *__restrict__ base, long long plane_stride, int32_t *dst,
                                      NaturalShape s, int planar) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= s.n_sites) return;
  dst[l] = base[site_addr_of(s, l, planar, plane_stride)];
}

// ---- OccLocation on the device (one chain): loc[cand][i] -> mol id, loc_size[cand],
// mol_loc[mol] -> position in its list.  Every site is a mutating site of one asymmetric
// unit orbit here, so mol id == l and candidate index == species index.
struct KLocation {
  int *loc;       // [K][n_sites]
  int *loc_size;  // [K]
  int *mol_loc;   // [n_sites]
};
// OccLocation::initialize (src/casm/monte/events/OccLocation.cc:68-114): lists are filled in
// site order, which fixes the order choose_mol draws from -- sequential by definition
__global__ void k_kstate_location_init(const uint8_t *__restrict__ nat, long long n_sites, int K, KLocation P) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int s = 0; s < K; ++s) P.loc_size[s] = 0;
  for (long long l = 0; l < n_sites; ++l) {
    const int s = nat[l];
    P.mol_loc[l] = P.loc_size[s];
    P.loc[(long long)s * n_sites + P.loc_size[s]++] = (int)l;
  }
}

struct KSerialArgs {
  uint8_t *nat;  // [chain][n_sites]
  NaturalShape shape;
  const KStateTables *tabs;
  MT64State *engines;
  unsigned long long *n_accept;
  KLocation loc;  // chain 0; chains are n_sites*(K+1)+K ints apart (see host)
  long long loc_chain_stride;  // ints between chains, for loc / mol_loc / loc_size separately scaled on the host
  long long n_passes;
};

// The reference's loop with the general proposal machinery, one thread per chain
// (thread 0 of the CTA; the CTA produces the Mersenne stream in blocks, see
// k_serial_reference).  Per step: choose_semigrand_canonical_swap draws
// random_real(total number of possible events), choose_mol draws
// random_int(size - 1) in the list of the swap's first candidate, and
// metropolis_acceptance draws random_real(1.0) only when dPhi >= 0.
__global__ void __launch_bounds__(kSerialThreads) k_kstate_serial(KSerialArgs A) {
  __shared__ unsigned long long mt[312], out[312];
  __shared__ int s_done;
  const int chain = blockIdx.x;
  const NaturalShape s = A.shape;
  uint8_t *nat = A.nat + (long long)chain * s.n_sites;
  const KStateTables *tab = A.tabs + chain;
  MT64State *eng = A.engines + chain;
  const int K = tab->K, z = tab->z;
  int *loc = A.loc.loc + (long long)chain * K * s.n_sites;
  int *loc_size = A.loc.loc_size + (long long)chain * kMaxSpecies;
  int *mol_loc = A.loc.mol_loc + (long long)chain * s.n_sites;

  for (int i = threadIdx.x; i < 312; i += blockDim.x) {
    mt[i] = eng->x[i];
    out[i] = mt64_temper(mt[i]);
  }
  if (threadIdx.x == 0) s_done = (A.n_passes <= 0);
  __syncthreads();
  int pos = eng->pos;
  long long pass = 0, step = 0, l = 0;
  unsigned long long n_acc = 0;
  int phase = 0, from = 0, to = 0, entry = 0;

  while (!s_done) {
    if (pos >= 312) {
      mt64_twist_block(mt);
      for (int i = threadIdx.x; i < 312; i += blockDim.x) out[i] = mt64_temper(mt[i]);
      pos = 0;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      bool done = false;
      while (pos < 312) {
        const unsigned long long u = out[pos++];
        if (phase == 0) {
          // choose_semigrand_canonical_swap (OccEventProposal.hh:262-306): swaps in the order of
          // make_semigrand_canonical_swaps (a outer, b inner, a != b); tsum accumulates cand_size(a)
          double total = 0.0;
          for (int a = 0; a < K; ++a)
            for (int b = 0; b < K; ++b)
              if (a != b) total = __dadd_rn(total, (double)loc_size[a]);
          double r = __dmul_rn(__ull2double_rn(u), 5.42101086242752217003726400434970855712890625e-20);
          if (r >= 1.0) r = 0.99999999999999988897769753748434595763683319091796875;
          const double rand = __dadd_rn(__dmul_rn(r, __dsub_rn(total, 0.0)), 0.0);
          double tsum = 0.0;
          from = -1;
          for (int a = 0; a < K && from < 0; ++a)
            for (int b = 0; b < K; ++b) {
              if (a == b) continue;
              tsum = __dadd_rn(tsum, (double)loc_size[a]);
              if (rand < tsum) {
                from = a;
                to = b;
                break;
              }
            }
          if (from < 0) {  // cannot happen (rand < total); keep the walk well-defined
            from = K - 1;
            to = K - 2;
          }
          phase = 1;
          continue;
        }
        if (phase == 1) {
          // choose_mol (OccLocation.hh:255-262): random_int(size - 1), libstdc++ Lemire with range = size
          const unsigned long long range = (unsigned long long)loc_size[from];
          const unsigned long long low = u * range;
          if (low < range && low < (0ull - range) % range) continue;  // redraw
          const long long pick = (long long)__umul64hi(u, range);
          l = loc[(long long)from * s.n_sites + pick];
          // neighbour configuration of site l
          const int i = (int)(l % s.n0);
          const long long rr = l / s.n0;
          const int j = (int)(rr % s.n1);
          const int k = (int)(rr / s.n1);
          const long long base = (long long)s.n0 * (j + (long long)s.n1 * k);
          const int ip = (i + 1 == s.n0) ? 0 : i + 1, im = (i == 0) ? s.n0 - 1 : i - 1;
          const int jp = (j + 1 == s.n1) ? 0 : j + 1, jm = (j == 0) ? s.n1 - 1 : j - 1;
          const long long kk = (long long)s.n1 * k;
          int cfg = kstate_weight(nat[base + ip], z) + kstate_weight(nat[base + im], z) +
                    kstate_weight(nat[i + (long long)s.n0 * (jp + kk)], z) + kstate_weight(nat[i + (long long)s.n0 * (jm + kk)], z);
          if (s.dim == 3) {
            const int kp = (k + 1 == s.n2) ? 0 : k + 1, km = (k == 0) ? s.n2 - 1 : k - 1;
            cfg += kstate_weight(nat[i + (long long)s.n0 * (j + (long long)s.n1 * kp)], z) +
                   kstate_weight(nat[i + (long long)s.n0 * (j + (long long)s.n1 * km)], z);
          }
          entry = kstate_entry(tab, from, to, cfg);
          if (!(tab->dPhi[entry] < 0.0)) {
            phase = 2;  // metropolis.hh:28-34: the uniform is drawn only when dPhi >= 0
            continue;
          }
        } else {
          double r = __dmul_rn(__ull2double_rn(u), 5.42101086242752217003726400434970855712890625e-20);
          if (r >= 1.0) r = 0.99999999999999988897769753748434595763683319091796875;
          r = __dadd_rn(__dmul_rn(r, __dsub_rn(1.0, 0.0)), 0.0);
          if (!(r < tab->prob[entry])) from = -1;  // rejected
        }
        if (from >= 0) {
          // OccLocation::apply (OccLocation.cc:263-282): swap-remove from the old list, append to the new
          nat[l] = (uint8_t)to;
          const int at = mol_loc[l];
          const int back = loc[(long long)from * s.n_sites + loc_size[from] - 1];
          loc[(long long)from * s.n_sites + at] = back;
          mol_loc[back] = at;
          loc_size[from]--;
          mol_loc[l] = loc_size[to];
          loc[(long long)to * s.n_sites + loc_size[to]++] = (int)l;
          ++n_acc;
        }
        phase = 0;
        if (++step == s.n_sites) {
          step = 0;
          if (++pass == A.n_passes) {
            done = true;
            break;
          }
        }
      }
      if (done) s_done = 1;
    }
    __syncthreads();
    if (!s_done) pos = 312;
  }
  if (threadIdx.x == 0) {
    eng->pos = pos;
    A.n_accept[chain] += n_acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 312; i += blockDim.x) eng->x[i] = mt[i];
}

}  // namespace cmg

// ===========================================================================
// N-fold way (rejection-free) driver for the Ising SGC model
// (include/casm/monte/methods/nfold.hh:80-147): per step the total rate, the
// selection of (event, time_increment), a sample if one is due -- taken BEFORE the
// event is applied, with the time increment as its weight -- and the event.  The
// event selector is outside the reference tree; this one is the Bortz-Kalos-
// Lebowitz selector over the 2*(2*dim+1) rate classes of the acceptance table
// (rate = 1 if dE < 0 else exp(-dE*beta)), with class lists kept like OccLocation
// keeps its candidate lists.  Sequential by nature: one thread per chain walks it on
// the reference's mt19937_64 stream (three draws per step), independent chains side
// by side.  It produces the weighted observations the statistics kernels accept.
// ===========================================================================
namespace cmg {

struct NfoldLists {
  int *members;     // [n_class][n_sites]
  int *n_members;   // [16]
  int *site_pos;    // [n_sites]
  uint8_t *site_class;  // [n_sites]
};
__device__ __forceinline__ void nfold_insert(const NfoldLists &P, long long n_sites, int l, int c) {
  P.site_class[l] = (uint8_t)c;
  P.site_pos[l] = P.n_members[c];
  P.members[(long long)c * n_sites + P.n_members[c]++] = l;
}
__device__ __forceinline__ void nfold_move(const NfoldLists &P, long long n_sites, int l, int c_new) {
  const int c = P.site_class[l];
  if (c == c_new) return;
  const int back = P.members[(long long)c * n_sites + P.n_members[c] - 1];
  P.members[(long long)c * n_sites + P.site_pos[l]] = back;
  P.site_pos[back] = P.site_pos[l];
  P.n_members[c]--;
  nfold_insert(P, n_sites, l, c_new);
}
// lists filled in site order (which fixes the order the member draw indexes)
__global__ void k_nfold_init(const uint8_t *__restrict__ nat, NaturalShape s, NfoldLists P) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int c = 0; c < 16; ++c) P.n_members[c] = 0;
  for (long long l = 0; l < s.n_sites; ++l)
    nfold_insert(P, s.n_sites, (int)l, 2 * natural_n_up32(nat, s.n0, s.n1, s.n2, s.dim, (unsigned int)l) + nat[l]);
}

struct NfoldArgs {
  uint8_t *nat;  // [chain][n_sites]
  NaturalShape shape;
  const ChainTables *tabs;
  MT64State *engines;
  NfoldLists lists;            // chain 0
  long long *cur_sb;           // [chain][2] current {ones, B}
  long long *series;           // slot of the first sample of this launch, chain 0: {ones, B}
  long long series_slot_stride;  // long long units between consecutive samples
  double *weight;              // [sample][chain] time increments, first sample of this launch
  double *rate_ratio;          // [sample][chain] total_rate / n_sites
  long long wstride;           // doubles between consecutive samples (= n_chains)
  double *time;                // [chain]
  long long n_steps;
  long long sample_period;     // steps; 0 = never
  long long step_base;         // steps done before this launch
};

__global__ void __launch_bounds__(kSerialThreads) k_nfold(NfoldArgs A) {
  __shared__ unsigned long long mt[312], out[312];
  __shared__ int s_done;
  const int chain = blockIdx.x;
  const NaturalShape s = A.shape;
  uint8_t *nat = A.nat + (long long)chain * s.n_sites;
  const ChainTables *tab = A.tabs + chain;
  MT64State *eng = A.engines + chain;
  const int z = 2 * s.dim, n_class = 2 * (z + 1);
  NfoldLists P;
  P.members = A.lists.members + (long long)chain * 16 * s.n_sites;
  P.n_members = A.lists.n_members + chain * 16;
  P.site_pos = A.lists.site_pos + (long long)chain * s.n_sites;
  P.site_class = A.lists.site_class + (long long)chain * s.n_sites;

  for (int i = threadIdx.x; i < 312; i += blockDim.x) {
    mt[i] = eng->x[i];
    out[i] = mt64_temper(mt[i]);
  }
  if (threadIdx.x == 0) s_done = (A.n_steps <= 0);
  __syncthreads();
  int pos = eng->pos;
  long long step = 0, slot = 0;
  long long ones = 0, B = 0;
  double time = 0.0, total_rate = 0.0;
  int phase = 0, chosen = 0, l = 0;
  double rate[16];
  if (threadIdx.x == 0) {
    ones = A.cur_sb[2 * chain];
    B = A.cur_sb[2 * chain + 1];
    time = A.time[chain];
    for (int c = 0; c < n_class; ++c) rate[c] = tab->dE[c] < 0.0 ? 1.0 : tab->prob[c];
  }

  while (!s_done) {
    if (pos >= 312) {
      mt64_twist_block(mt);
      for (int i = threadIdx.x; i < 312; i += blockDim.x) out[i] = mt64_temper(mt[i]);
      pos = 0;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      bool done = false;
      while (pos < 312) {
        const unsigned long long u = out[pos++];
        if (phase == 0) {
          // total rate (before selection) and the class of the event
          total_rate = 0.0;
          for (int c = 0; c < n_class; ++c) total_rate = __dadd_rn(total_rate, __dmul_rn((double)P.n_members[c], rate[c]));
          double r = __dmul_rn(__ull2double_rn(u), 5.42101086242752217003726400434970855712890625e-20);
          if (r >= 1.0) r = 0.99999999999999988897769753748434595763683319091796875;
          const double u1 = __dadd_rn(__dmul_rn(r, __dsub_rn(total_rate, 0.0)), 0.0);
          chosen = -1;
          double cum = 0.0;
          for (int c = 0; c < n_class; ++c) {
            cum = __dadd_rn(cum, __dmul_rn((double)P.n_members[c], rate[c]));
            if (u1 < cum) {
              chosen = c;
              break;
            }
          }
          if (chosen < 0)
            for (int c = n_class - 1; c >= 0; --c)
              if (P.n_members[c] > 0) {
                chosen = c;
                break;
              }
          phase = 1;
          continue;
        }
        if (phase == 1) {
          const unsigned long long range = (unsigned long long)P.n_members[chosen];
          const unsigned long long low = u * range;
          if (low < range && low < (0ull - range) % range) continue;  // Lemire redraw
          l = P.members[(long long)chosen * s.n_sites + (long long)__umul64hi(u, range)];
          phase = 2;
          continue;
        }
        // phase 2: the time increment, the sample if one is due, the event
        double r = __dmul_rn(__ull2double_rn(u), 5.42101086242752217003726400434970855712890625e-20);
        if (r >= 1.0) r = 0.99999999999999988897769753748434595763683319091796875;
        r = __dadd_rn(__dmul_rn(r, __dsub_rn(1.0, 0.0)), 0.0);
        const double time_increment = __ddiv_rn(-log(__dsub_rn(1.0, r)), total_rate);
        if (A.sample_period > 0 && ((A.step_base + step + 1) % A.sample_period) == 0) {
          long long *dst = A.series + slot * A.series_slot_stride + 2 * chain;
          dst[0] = ones;
          dst[1] = B;
          A.weight[slot * A.wstride + chain] = time_increment;
          A.rate_ratio[slot * A.wstride + chain] = __ddiv_rn(total_rate, (double)s.n_sites);
          ++slot;
        }
        {
          const int n_up = natural_n_up32(nat, s.n0, s.n1, s.n2, s.dim, (unsigned int)l);
          const int b = nat[l];
          const int sgn = 2 * b - 1;
          B += (long long)(-2 * sgn) * (2 * n_up - z);
          ones += b ? -1 : 1;
          nat[l] = (uint8_t)(b ^ 1);
          nfold_move(P, s.n_sites, l, 2 * n_up + (b ^ 1));
          // the neighbours change class: +i, +j, -i, -j [, +k, -k]
          const unsigned int ul = (unsigned int)l;
          const unsigned int rr = ul / (unsigned int)s.n0;
          const int i = (int)(ul - rr * (unsigned int)s.n0);
          int j = (int)rr, k = 0;
          if (s.dim == 3) {
            k = (int)(rr / (unsigned int)s.n1);
            j = (int)(rr - (unsigned int)k * (unsigned int)s.n1);
          }
          const int ip = (i + 1 == s.n0) ? 0 : i + 1, im = (i == 0) ? s.n0 - 1 : i - 1;
          const int jp = (j + 1 == s.n1) ? 0 : j + 1, jm = (j == 0) ? s.n1 - 1 : j - 1;
          int nb[6];
          nb[0] = ip + s.n0 * (j + s.n1 * k);
          nb[1] = i + s.n0 * (jp + s.n1 * k);
          nb[2] = im + s.n0 * (j + s.n1 * k);
          nb[3] = i + s.n0 * (jm + s.n1 * k);
          if (s.dim == 3) {
            const int kp = (k + 1 == s.n2) ? 0 : k + 1, km = (k == 0) ? s.n2 - 1 : k - 1;
            nb[4] = i + s.n0 * (j + s.n1 * kp);
            nb[5] = i + s.n0 * (j + s.n1 * km);
          }
          for (int d = 0; d < z; ++d)
            nfold_move(P, s.n_sites, nb[d],
                       2 * natural_n_up32(nat, s.n0, s.n1, s.n2, s.dim, (unsigned int)nb[d]) + nat[nb[d]]);
        }
        time = __dadd_rn(time, time_increment);
        phase = 0;
        if (++step == A.n_steps) {
          done = true;
          break;
        }
      }
      if (done) s_done = 1;
    }
    __syncthreads();
    if (!s_done) pos = 312;
  }
  if (threadIdx.x == 0) {
    eng->pos = pos;
    A.cur_sb[2 * chain] = ones;
    A.cur_sb[2 * chain + 1] = B;
    A.time[chain] = time;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 312; i += blockDim.x) eng->x[i] = mt[i];
}

}  // namespace cmg
