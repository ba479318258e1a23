// common.cuh — device parameter block, exact cell arithmetic, pair laws and the stencil walker shared by all
// kernels of libmavi_cuda.so.  sm_100a only.  Reference citations are relative to /root/reference/.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mavi.h"
#include "real.cuh"

namespace MAVI_NS {

// One Line2D (src/configs.jl:95-117) with the frame the ctor derives, laid out as 9 doubles on the device.
struct DevLine {
  real p1[2], p2[2], normal[2], tangent[2], length;
};

struct DevSpace {
  int wall, geom;
  real rect_bl[2], rect_sz[2];
  real cc[2], cr;
  const DevLine *lines;
  int n_lines;
  int pot_kind;
  real pot[4];
  real pot_cut2;  // largest r2 with sqrt(r2) <= dist_max (exact cutoff test without sqrt)
  int pot_mode;
  int n_pot_types;                      // PotentialVector: pot_t[type] instead of pot (Mavi.Rings: the ring type)
  real pot_t[MAVI_MAX_POT_TYPES][4];
};

struct DevRings {
  int num_types, n_max;
  int num_rings;
  const real *p0, *relax_time, *vo, *mobility, *rot_diff, *k_area, *k_spring, *l_spring;
  const int *num_particles;
  const real *interaction;  // [t1][t2][7]: k_rep,k_atr,dist_eq,dist_max, cut2, eq2_lo (exact r2 thresholds), 1/dist_eq
  const int *types;           // 0-based per ring, or nullptr
};

// Kernel parameter block (passed by value as __grid_constant__).
// Split launches of the slab step: the edge blocks (blk_mode 2) are the MAVI_EDGE_COLS owned columns next to each halo — the
// only blocks that read a halo column, and the only ones an emigrant can come from unless it crosses >= MAVI_EDGE_COLS
// columns in one step (reported loudly) — ~3 % of the particles of a 100-column slab, so the boundary launch is short.
#define MAVI_EDGE_COLS 3

struct DevParams {
  int n;               // particle slots held by this device
  int n_count;         // get_num_total_particles(state): number of active ids
  int num_cols, num_rows, num_cells;
  int wrap_cols, wrap_rows;  // stencil index wrap (periodic main wall)
  // Padded tile layout: a tile = MAVI_TR consecutive cell rows of ONE cell column with `cap` particle slots; tile t owns
  // slots [t*cap, (t+1)*cap), particles sorted by (cell, id) in its prefix.  Inactive slots live in a tail region.
  int tpc;        // tiles per column = ceil(num_rows / MAVI_TR)
  int nt;         // number of tiles = num_cols * tpc
  int cap;        // slots per tile
  int n_active;   // particles living in tiles; ranks >= n_active are the inactive tail
  int tail_base;  // first slot of the inactive tail = nt * cap
  int inbox_cap;  // per-tile capacity for particles arriving from other tiles in one step
  int mv_cap;     // capacity of the per-step inter-tile mover list
  int chg_cap;    // capacity of the per-step changed-cell list (force carry)
  int blk_cols, blk_per_row;  // tile-block force kernels: own tiles (columns) per CTA, CTAs per tile row
  int pipe_items; // pipelined kernels: tile blocks a CTA takes before it exits (0: persistent, until the work counter runs out)
  int pipe_reserved;  // pipelined kernels: CTAs that land on an SM with %smid < pipe_reserved exit at once (slab step, grid_pipe())
  int blk_mode;   // 0: all blocks; 1: the interior of every tile row (all owned columns but MAVI_EDGE_COLS at each end, in blocks
                  // of blk_cols); 2: the two edge blocks of every tile row (MAVI_EDGE_COLS columns each) — blk_items_per_row()
  // x-slab domain decomposition (one process per GPU): the local grid is [left halo | owned columns | right halo];
  // cell arithmetic stays GLOBAL (bit-exact global cell ids), only the column index is shifted into the local frame.
  int slab;                   // 1 = slab mode
  int gcols;                  // global number of cell columns (== num_cols when slab == 0)
  int col_lo;                 // global column of local column 1 (slab mode)
  int ord_cols, ord_col0;     // columns enumerated by the rank order: owned columns only
  int nt_ord;                 // ord_cols * tpc
  int seam_left, seam_right;  // the halo column lies across the periodic seam -> minimum image for pairs through it
  unsigned int rows_mul, rows_shr;  // magic number division by num_rows
  unsigned int tpc_mul, tpc_shr;    // ... by tpc
  unsigned int cols_mul, cols_shr;  // ... by ord_cols
  int periodic;              // calc_diff applies the minimum image (src/integration.jl:43-48)
  double grid_bl[2], grid_h, cl, ch;  // Float64 in BOTH builds: Chunks keeps chunk_length / chunk_height and the geometry in
                                      // Float64 (src/chunks.jl:13,27-30), so a Float32 state is binned with Float64 arithmetic
  real size[2], half[2];   // main rectangle size and size/2
  int dynamics;
  real dyn[8];
  // derived pair-law constants
  real lj_sig2, lj_24eps;        // LJ / RTP:  F/d = 24 eps s6 (2 s6 - 1) / r2,  s6 = (sig^2/r2)^3
  real lj_c48, lj_c24;           // 48 eps / sig^2, 24 eps / sig^2:  F/d = u^4 (c48 u^3 - c24),  u = sig^2 / r2
  int fast_interior;               // grid >= 8x8: interior cells may skip the minimum image (guarded, see kernels.cu)
  real cut2;                     // exact r2 threshold of the law's cutoff (HarmTrunc dist_max, Szabo r_max, RTP 2^(1/6) sigma)
  real eq2_lo;                   // HarmTrunc: smallest r2 with sqrt(r2) >= dist_eq  (d < dist_eq  <=>  r2 < eq2_lo)
  real szabo_eq2_hi;             // Szabo: largest r2 with sqrt(r2) <= r_eq          (d > r_eq     <=>  r2 > szabo_eq2_hi)
  real harm_inv_deq;             // HarmTrunc: 1/dist_eq
  real szabo_fadh, szabo_frep;   // Szabo: k_adh/r_eq, k_rep/(r_max-r_eq)  (src/integration.jl:79-83)
  real szabo_inv_tau, szabo_namp;  // Szabo: 1/relax_time, sqrt(2 rot_diff dt)  (src/integration.jl:460), hoisted out of the kernel
  real particle_radius;
  real dt, term, hdt;            // dt, dt^2/2, dt/2  (src/integration.jl:420-430)
  int n_spaces;
  int has_force_walls;
  int wall_fast;             // 1: the space is ONE periodic rectangle (walls! fast path)
  real wall_ctr[2];        // its centre bl + size/2
  DevSpace spaces[MAVI_MAX_SPACES];
  int rng_mode;
  unsigned long long seed;
  DevRings rings;
};

// tile blocks of one tile row for the current blk_mode, and the owned columns [c_begin, c_end) of block `bcol`
__host__ __device__ inline int blk_items_per_row(const DevParams &p) {
  if (p.blk_mode == 0) return p.blk_per_row;
  if (p.blk_mode == 1) return (p.ord_cols - 2 * MAVI_EDGE_COLS + p.blk_cols - 1) / p.blk_cols;
  return 2;
}
__host__ __device__ inline void blk_columns(const DevParams &p, int bcol, int &c_begin, int &c_end) {
  const int c0 = p.ord_col0, c1 = p.ord_col0 + p.ord_cols;
  if (p.blk_mode == 0) {
    c_begin = c0 + bcol * p.blk_cols;
    c_end = c_begin + p.blk_cols < c1 ? c_begin + p.blk_cols : c1;
  } else if (p.blk_mode == 1) {
    c_begin = c0 + MAVI_EDGE_COLS + bcol * p.blk_cols;
    c_end = c_begin + p.blk_cols < c1 - MAVI_EDGE_COLS ? c_begin + p.blk_cols : c1 - MAVI_EDGE_COLS;
  } else {
    c_begin = bcol ? c1 - MAVI_EDGE_COLS : c0;
    c_end = c_begin + MAVI_EDGE_COLS;
  }
}


enum { ERRBIT_OUT_OF_GRID = 1, ERRBIT_NAN = 2, ERRBIT_OUTSIDE_SPACE = 4, ERRBIT_OOG_PENDING = 8 };

// flags[] layout (device control word, mirrored to pinned host memory once per step)
enum {
  FLAG_ERR = 0,
  // reset at the start of every step:
  FLAG_CHANGED = 1, FLAG_BIGMOVE = 2, FLAG_NFIX = 3, FLAG_NMV = 4, FLAG_RAN = 5,
  FLAG_PER_STEP = 5,  // number of per-step words starting at index 1
  // latched until handled by the host:
  FLAG_OVERFLOW = 6, FLAG_MAXCOUNT = 7, FLAG_TAIL = 8, FLAG_MAXINBOX = 9, FLAG_MAXINBOX_TILE = 10,
  FLAG_STEPS = 11,    // steps that really ran since the last upload (device-side step counter)
  FLAG_SCRATCH = 12,
  // force carry (see k_newton_b): cells whose membership changed in this step, and the big-drift guard of the NEXT step
  FLAG_NCHG = 16, FLAG_BIGMOVE_NEXT = 17,
  // slab mode: emigrants of this step towards the left / right neighbour, and the counts received from them
  FLAG_NEM0 = 18, FLAG_NEM1 = 19, FLAG_NEMR0 = 20, FLAG_NEMR1 = 21,
  // instrumentation (mavi_counters): changed-cell records and inter-tile movers summed over the steps since the last upload
  FLAG_CUM_CHG = 22, FLAG_CUM_MV = 23, FLAG_CUM_DIRTY = 13, FLAG_CUM_EM = 14,
  FLAG_NMOVED = 15,  // per step: particles whose cell changed (re-binned by the incremental repair)
  // persistent tile-block kernels: next work item of the launch on the main stream [0] / of the boundary-block launch [1]
  FLAG_WORK0 = 24, FLAG_WORK1 = 25,
  FLAG_INBOX_STEP = 26,  // per step: largest inbox population (with FLAG_MAXCOUNT it tells whether a tile CAN overflow at all)
  FLAG_COUNT = 32
};

// A step whose tile repair overflowed (or that pushed a particle out of the grid) leaves the layout un-repaired: every
// kernel of the FOLLOWING steps sees the latched flag and returns at once, so the host may enqueue many steps and look
// at the control words only occasionally; FLAG_STEPS counts the steps that really ran.
__device__ __forceinline__ bool step_poisoned(const int *flags) {
  return (flags[FLAG_OVERFLOW] != 0) || ((flags[FLAG_ERR] & ERRBIT_OOG_PENDING) != 0);
}

// raise a high-water mark without hammering one address with atomics: almost every caller sees a value that is already there
__device__ __forceinline__ void raise_mark(int *word, int v) {
  if (v > *reinterpret_cast<volatile int *>(word)) atomicMax(word, v);
}

#define MAVI_TR 32  // cell rows per tile


// slab mode: record of a particle that leaves this rank's columns (written by the integrate kernel, shipped as is)
struct EmRec {
  real2 pos, second, force;
  unsigned int idflag;
  int pad[3];
};

// index of (tile, local row) in tstart[]: tstart[tile*(MAVI_TR+1) + lr] = first slot of that cell's particles,
// entry MAVI_TR = end of the tile's particles.
// x / d for 0 <= x < 2^31 with a host-computed (mul, shr) pair: q = umulhi(x, mul) >> shr  (Granlund-Montgomery)
__device__ __forceinline__ int fastdiv(int x, unsigned int mul, unsigned int shr) {
  return mul ? (int)(__umulhi((unsigned int)x, mul) >> shr) : x;  // mul == 0 encodes d == 1
}
__device__ __forceinline__ int div_rows(const DevParams &p, int x) { return fastdiv(x, p.rows_mul, p.rows_shr); }
__device__ __forceinline__ int div_tpc(const DevParams &p, int x) { return fastdiv(x, p.tpc_mul, p.tpc_shr); }
__device__ __forceinline__ int div_cols(const DevParams &p, int x) { return fastdiv(x, p.cols_mul, p.cols_shr); }

__device__ __forceinline__ int tile_of_cell(const DevParams &p, int cell) {
  const int col = div_rows(p, cell), row = cell - col * p.num_rows;
  return col * p.tpc + row / MAVI_TR;
}
// Kernels enumerate tiles TILE-ROW-MAJOR (order index o = tile_row * num_cols + col): 256 consecutive ranks are then a
// block of ~6 adjacent columns x 32 rows whose neighbours are mostly the block's own particles (L1 hits).
__device__ __forceinline__ int tile_of_order(const DevParams &p, int o) {
  const int tr = div_cols(p, o), col = o - tr * p.ord_cols + p.ord_col0;
  return col * p.tpc + tr;
}
__device__ __forceinline__ int tq_of(const DevParams &p, int col, int row) {
  const int tr = row / MAVI_TR;
  return (col * p.tpc + tr) * (MAVI_TR + 1) + (row - tr * MAVI_TR);
}

#define MAVI_INACTIVE_BIT 0x80000000u

// ---------------------------------------------------------------------------------------------------------
// Exact Base.div(x::Float64, y::Float64) = round((x - rem(x,y))/y) for y > 0 (call sites src/chunks.jl:129-130).
// rem is exact, so the reference value is trunc(x/y) of the REAL quotient.  fl(x/y) can be off by one unit when x
// is a rounded multiple of y; one FMA gives the sign of the exact remainder and fixes it.  Verified against the
// fmod formulation (oracle mor_julia_div) in tests/test_oracle_kat.py and tests/test_gpu_core.py (test_cell_assignment_*).
__device__ __forceinline__ double julia_div_pos(double x, double y) {
  double ax = fabs(x);
  double q = trunc(ax / y);
  double rem = fma(-q, y, ax);  // sign (and zero-ness) of ax - q*y is exact
  if (rem < 0.0) q -= 1.0;
  else if (rem >= y) q += 1.0;
  return copysign(q, x);
}

// update_particle_chunk! (src/chunks.jl:120-147): 0-based linear cell id (row fastest), or -1 if out of grid.
// Positions of a Float32 state are promoted: div(-pos[2] + bottom_left[2] + space_h, chunk_h) mixes Float32 and Float64.
__device__ __forceinline__ int cell_of_point(const DevParams &p, double x, double y) {
  double rowf = julia_div_pos(-y + p.grid_bl[1] + p.grid_h, p.ch);
  double colf = julia_div_pos(x - p.grid_bl[0], p.cl);
  if (!(fabs(rowf) < 2.0e9) || !(fabs(colf) < 2.0e9)) return -1;  // NaN/Inf -> InexactError in the reference
  int row = (int)rowf + 1, col = (int)colf + 1;
  row -= (row == p.num_rows + 1) ? 1 : 0;
  col -= (col == p.gcols + 1) ? 1 : 0;
  if (row < 1 || row > p.num_rows || col < 1 || col > p.gcols) return -1;
  int lcol = col - 1;
  if (p.slab) {  // global column -> local frame [0 = left halo, 1..m owned, m+1 = right halo], periodic in x
    int d = lcol - p.col_lo;
    if (d < 0) d += p.gcols;  // [0, gcols): columns to the right of my first one
    lcol = d + 1;
    if (lcol > p.num_cols - 1) {  // beyond the right halo: the left halo is the only other column I hold
      lcol -= p.gcols;
      if (lcol != 0) return -1;  // more than one column beyond the slab in one step
    }
  }
  return (row - 1) + p.num_rows * lcol;
}

// Exact test "update_particle_chunk! would put (x, y) into `cell` again" without a division: the reference index is
// trunc of the REAL quotient t/c, so index == k  <=>  k*c <= t < (k+1)*c with the products taken exactly; for a double t
// that is  RU(k*c) <= t < RU((k+1)*c)  (RU = round-up multiply).  Edge conventions of src/chunks.jl:129-142: index -0
// (t in (-c, 0)) maps to the first cell, index n (t in [n*c, (n+1)*c)) is clamped to the last cell.
__device__ __forceinline__ bool axis_in_cell(double t, int k, int n, double c) {
  bool lo = (k == 0) ? (t > -c) : (t >= __dmul_ru((double)k, c));
  bool hi = t < __dmul_ru((double)((k == n - 1) ? n + 1 : k + 1), c);
  return lo && hi;
}
__device__ __forceinline__ bool still_in_cell(const DevParams &p, double x, double y, int cell) {
  int col = div_rows(p, cell);
  const int row = cell - col * p.num_rows;
  if (p.slab) {  // local -> global column
    col = col + p.col_lo - 1;
    if (col < 0) col += p.gcols;
    else if (col >= p.gcols) col -= p.gcols;
  }
  return axis_in_cell(-y + p.grid_bl[1] + p.grid_h, row, p.num_rows, p.ch) &&
         axis_in_cell(x - p.grid_bl[0], col, p.gcols, p.cl);
}

// calc_diff component (src/integration.jl:38-48): strict '>', one image.
template <bool PERIODIC>
__device__ __forceinline__ real min_image(real d, real half, real size) {
  if (PERIODIC) {
    if (fabs(d) > half) d -= copysign(size, d);
  }
  return d;
}

// r2 exactly as the reference rounds it: sum(dr.^2) = fl(fl(dx*dx) + fl(dy*dy)), no contraction.
__device__ __forceinline__ real dist2_exact(real dx, real dy) {
  return add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
}

// ---------------------------------------------------------------------------------------------------------
// Pair laws: coefficient c with force_on_i = c * dr  (dr = r_i - r_j, minimum image applied).
template <int DYN>
__device__ __forceinline__ real pair_coef(const DevParams &p, real r2);

// LenJonesCfg, src/configs.jl:389-397: fmod/d = 4 eps (12 sig^12/d^14 - 6 sig^6/d^8); no cutoff.
template <>
__device__ __forceinline__ real pair_coef<MAVI_DYN_LJ>(const DevParams &p, real r2) {
  real u = p.lj_sig2 * fast_rcp(r2);
  real u2 = u * u;
  real u3 = u2 * u;
  real u4 = u2 * u2;
  return u4 * fma(u3, p.lj_c48, -p.lj_c24);  // 6 FP64 ops + 3 for the reciprocal
}

// HarmTruncCfg, src/configs.jl:354-368: 0 beyond dist_max; fmod/d = -k (d/d_eq - 1)/d = k (1/d - 1/d_eq).
template <>
__device__ __forceinline__ real pair_coef<MAVI_DYN_HARMTRUNC>(const DevParams &p, real r2) {
  if (r2 > p.cut2) return 0.0;
  real k = (r2 < p.eq2_lo) ? p.dyn[0] : p.dyn[1];
  real inv_d = rsqrt(r2);
  return k * (inv_d - p.harm_inv_deq);
}

// SzaboCfg, src/integration.jl:68-87: -f_mod (d - r_eq) * dr with dr NOT normalised (kept).
template <>
__device__ __forceinline__ real pair_coef<MAVI_DYN_SZABO>(const DevParams &p, real r2) {
  if (r2 > p.cut2) return 0.0;
  real d = sqrt(r2);
  real f_mod = (r2 > p.szabo_eq2_hi) ? p.szabo_fadh : p.szabo_frep;
  return -f_mod * (d - p.dyn[5]);
}

// RunTumbleCfg, src/integration.jl:89-109: WCA, cutoff 2^(1/6) sigma.
template <>
__device__ __forceinline__ real pair_coef<MAVI_DYN_RTP>(const DevParams &p, real r2) {
  if (r2 > p.cut2) return 0.0;
  real u = p.lj_sig2 * fast_rcp(r2);
  real u2 = u * u;
  real u3 = u2 * u;
  real u4 = u2 * u2;
  return u4 * fma(u3, p.lj_c48, -p.lj_c24);
}

// ---------------------------------------------------------------------------------------------------------
// Stencil walker.  The reference enumerates same-cell pairs (j > i) plus the half stencil of each cell
// (src/chunks.jl:61-118, src/integration.jl:116-156) and scatters +f/-f.  As a gather, particle i sees every
// particle of its own cell and of the half stencil UNITED WITH ITS MIRROR IMAGE = the 8 surrounding cells
// (wrapped when the main wall is periodic, clipped otherwise).  Visiting wrapped rows/columns individually keeps
// the reference's double counting on 2-row / 2-column periodic grids.  This generic walker visits the 9 cells one by
// one through tstart[]; interior cells away from tile edges use the flat 3-run fast path in kernels.cu instead.
//   f(j) is called for every neighbour slot j != self.
template <typename F>
__device__ __forceinline__ void for_each_neighbor(const DevParams &p, const int *__restrict__ tstart, int cell, int self,
                                                  F &&f) {
  const int R = p.num_rows, Cn = p.num_cols;
  const int col = cell / R, row = cell - col * R;
#pragma unroll 1
  for (int dc = -1; dc <= 1; dc++) {
    int c2 = col + dc;
    if (c2 < 0) {
      if (!p.wrap_cols) continue;
      c2 = Cn - 1;
    } else if (c2 >= Cn) {
      if (!p.wrap_cols) continue;
      c2 = 0;
    }
#pragma unroll 1
    for (int dr = -1; dr <= 1; dr++) {
      int r2 = row + dr;
      if (r2 < 0) {
        if (!p.wrap_rows) continue;
        r2 = R - 1;
      } else if (r2 >= R) {
        if (!p.wrap_rows) continue;
        r2 = 0;
      }
      const int q = tq_of(p, c2, r2);
      int jb = __ldg(tstart + q), je = __ldg(tstart + q + 1);
      for (int j = jb; j < je; j++)
        if (j != self) f(j);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (production mode; the reference's Julia RNG streams are not reproducible,
// SURVEY.md 7 "RNG parity").  key = seed, counter = (particle/ring id, step, stream).
__device__ __forceinline__ void philox4x32(unsigned int c[4], unsigned int k0, unsigned int k1) {
#pragma unroll
  for (int i = 0; i < 10; i++) {
    unsigned int hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    unsigned int hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    unsigned int n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__device__ __forceinline__ double u01_from_bits(unsigned int hi, unsigned int lo) {
  unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return (double)(v >> 11) * (1.0 / 9007199254740992.0);  // [0,1)
}

// two uniforms in [0,1) for (id, step)
__device__ __forceinline__ void philox_uniform2(unsigned long long seed, unsigned int id, unsigned long long step,
                                                double &u0, double &u1) {
  unsigned int c[4] = {id, (unsigned int)step, (unsigned int)(step >> 32), 0x4d415649u};
  philox4x32(c, (unsigned int)seed, (unsigned int)(seed >> 32));
  u0 = u01_from_bits(c[0], c[1]);
  u1 = u01_from_bits(c[2], c[3]);
}

// Philox2x32-10 (Salmon et al., same construction with one multiplier): 64-bit counter, 32-bit key — half the multiplies of
// the 4x32 generator for the draws that need only two words.
__device__ __forceinline__ void philox2x32(unsigned int &c0, unsigned int &c1, unsigned int k) {
#pragma unroll
  for (int i = 0; i < 10; i++) {
    const unsigned int hi = __umulhi(0xD256D193u, c0), lo = 0xD256D193u * c0;
    c0 = hi ^ k ^ c1;
    c1 = lo;
    k += 0x9E3779B9u;
  }
}

// one standard normal (Box-Muller) for (id, step).  Production-mode noise only has to be N(0,1) to statistical accuracy
// (the reference draws from Julia's ziggurat randn, which no device stream can reproduce), so the transform runs in single
// precision with the hardware log / cos: 24-bit uniforms, u0 in (0,1) -> |z| <= sqrt(2 ln 2^25) = 5.9.  counter = (id, low
// step word), key = seed and high step word folded into 32 bits.  (The Szabo kernel spent 19 % of its instructions in the
// double-precision transform in round 1, profiles/r01_ncu_szabo_rings.md, and ~11 % in Philox4x32 + logf / cospif after that.)
__device__ __forceinline__ double philox_normal(unsigned long long seed, unsigned int id, unsigned long long step) {
  unsigned int c0 = id, c1 = (unsigned int)step;
  const unsigned int key = (unsigned int)seed ^ ((unsigned int)(seed >> 32) * 0x9E3779B9u) ^
                           ((unsigned int)(step >> 32) * 0x85EBCA6Bu) ^ 0x4d415649u;
  philox2x32(c0, c1, key);
  const float u0 = ((float)(c0 >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  const float u1 = (float)(c1 >> 8) * (1.0f / 16777216.0f);           // [0,1)
  const float r = sqrtf(-2.0f * __logf(u0));
  return (double)(r * __cosf(6.2831853f * u1));
}

__device__ __forceinline__ real sign_d(real x) { return x > real(0) ? real(1) : (x < real(0) ? real(-1) : x); }

}  // namespace MAVI_NS
