// kernels.cu — cell binning (counting sort + prefix scan + stable physical re-order), fused force+integrate
// passes, energy reductions.  sm_100a.  No tensor cores: the path is FP64 vector math over HBM-resident SoA arrays
// (double2 loads), not a dense contraction.  Reference citations are relative to /root/reference/.
#include "kernels.cuh"
#include "walls.cuh"

namespace mavi {

constexpr int TPB = 256;
static inline int nblk(long long n, int tpb = TPB) { return (int)((n + tpb - 1) / tpb); }

#define MAVI_LAUNCH(ctx, kernel, grid, block, ...)              \
  do {                                                          \
    kernel<<<(grid), (block), 0, (ctx).stream>>>(__VA_ARGS__);  \
    (*(ctx).launches)++;                                        \
  } while (0)

enum { ERRBIT_OUTSIDE_SPACE = 4 };

// =========================================================================================================
// Binning: update_chunks! (src/chunks.jl:150-163, src/integration.jl:54-59) as a counting sort.
// =========================================================================================================

// cell id per slot + histogram; also counts how many particles left the cell they are currently sorted under.
__global__ void k_cell_index(const __grid_constant__ DevParams p, const double2 *__restrict__ pos,
                             const unsigned int *__restrict__ idflag, const int *__restrict__ cell_old,
                             int *__restrict__ cell_new, int *__restrict__ count, int *__restrict__ flags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  bool changed = false;
  if (k < p.n) {
    int c;
    if (idflag[k] & MAVI_INACTIVE_BIT) {
      c = p.num_cells;  // pseudo-cell of inactive slots (never binned by the reference: active ids only)
    } else {
      double2 r = pos[k];
      c = cell_of_point(p, r.x, r.y);
      if (c < 0) {  // BoundsError in the reference (src/chunks.jl:144-146)
        atomicOr(&flags[0], ERRBIT_OUT_OF_GRID);
        c = 0;
      }
    }
    cell_new[k] = c;
    atomicAdd(&count[c], 1);
    changed = (c != cell_old[k]);
  }
  unsigned int m = __ballot_sync(0xffffffffu, changed);
  if (m && (threadIdx.x & 31) == 0) atomicAdd(&flags[1], __popc(m));
}

void launch_cell_index(const LaunchCtx &c, const DevParams &p, const double2 *pos, const unsigned int *idflag,
                       const int *cell_old, int *cell_new, int *count, int *flags) {
  MAVI_LAUNCH(c, k_cell_index, nblk(p.n), TPB, p, pos, idflag, cell_old, cell_new, count, flags);
}

// ---- exclusive prefix scan (reduce / top / final), 4096 items per block ------------------------------------
constexpr int SCAN_TPB = 256, SCAN_ITEMS = 16, SCAN_BLOCK = SCAN_TPB * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
  __shared__ int warp_sums[SCAN_TPB / 32];
  __shared__ int s_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  if (w == 0) {
    int ws = lane < SCAN_TPB / 32 ? warp_sums[lane] : 0;
    int wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < SCAN_TPB / 32) warp_sums[lane] = wi - ws;
    if (lane == SCAN_TPB / 32 - 1) s_total = wi;
  }
  __syncthreads();
  if (total) *total = s_total;
  return incl - v + warp_sums[w];
}

__global__ void k_scan_reduce(const int *__restrict__ in, int *__restrict__ partials, int n) {
  int base = blockIdx.x * SCAN_BLOCK;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int idx = base + i * SCAN_TPB + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  int tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

__global__ void k_scan_top(int *__restrict__ partials, int nb) {
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_TPB) {
    int idx = base + threadIdx.x;
    int v = idx < nb ? partials[idx] : 0;
    int tot;
    int ex = block_exclusive_scan(v, &tot);
    int carry = carry_s;
    if (idx < nb) partials[idx] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
}

__global__ void k_scan_final(const int *__restrict__ in, int *__restrict__ out, const int *__restrict__ partials, int n) {
  int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int idx = base + i;
    v[i] = idx < n ? in[idx] : 0;
    s += v[i];
  }
  int ex = block_exclusive_scan(s, nullptr) + partials[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int idx = base + i;
    if (idx < n) out[idx] = ex;
    ex += v[i];
  }
}

void launch_exclusive_scan(const LaunchCtx &c, const int *in, int *out, int *partials, int n) {
  int nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  MAVI_LAUNCH(c, k_scan_reduce, nb, SCAN_TPB, in, partials, n);
  MAVI_LAUNCH(c, k_scan_top, 1, SCAN_TPB, partials, nb);
  MAVI_LAUNCH(c, k_scan_final, nb, SCAN_TPB, in, out, partials, n);
}

// scatter slot ids into their cell range (cursor = count, consumed down to zero)
__global__ void k_scatter(const __grid_constant__ DevParams p, const int *__restrict__ cell_new,
                          const int *__restrict__ start, int *__restrict__ count, int *__restrict__ perm) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n) return;
  int c = cell_new[k];
  int slot = start[c] + atomicSub(&count[c], 1) - 1;
  perm[slot] = k;
}

void launch_scatter(const LaunchCtx &c, const DevParams &p, const int *cell_new, const int *start, int *count, int *perm) {
  MAVI_LAUNCH(c, k_scatter, nblk(p.n), TPB, p, cell_new, start, count, perm);
}

// Stable placement + physical re-order: the particle scattered to slot s goes to start[cell] + (rank of its original
// id inside the cell), i.e. ascending ids per cell like the reference's fill loop (src/chunks.jl:153-155).  Makes the
// layout (and every force summation order) independent of atomic scheduling -> bit-reproducible runs.
__global__ void k_gather(const __grid_constant__ DevParams p, const int *__restrict__ perm,
                         const int *__restrict__ cell_new, const int *__restrict__ start,
                         const double2 *__restrict__ pos_s, double2 *__restrict__ pos_d,
                         const double2 *__restrict__ vel_s, double2 *__restrict__ vel_d,
                         const double *__restrict__ ang_s, double *__restrict__ ang_d,
                         const unsigned int *__restrict__ id_s, unsigned int *__restrict__ id_d,
                         int *__restrict__ cell_d, const double2 *__restrict__ f_s, double2 *__restrict__ f_d) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.n) return;
  int src = perm[s];
  int c = cell_new[src];
  unsigned int idf = id_s[src];
  int d = s;
  if (c < p.num_cells) {
    int b = start[c], e = start[c + 1];
    unsigned int myid = idf & ~MAVI_INACTIVE_BIT;
    int rank = 0;
    for (int t = b; t < e; t++) {
      if (t == s) continue;
      unsigned int other = id_s[perm[t]] & ~MAVI_INACTIVE_BIT;
      rank += other < myid;
    }
    d = b + rank;
  }
  pos_d[d] = pos_s[src];
  if (vel_s) vel_d[d] = vel_s[src];
  if (ang_s) ang_d[d] = ang_s[src];
  id_d[d] = idf;
  cell_d[d] = c;
  if (f_s) f_d[d] = f_s[src];
}

void launch_gather(const LaunchCtx &c, const DevParams &p, const int *perm, const int *cell_new, const int *start,
                   const DevArrays &a, int src, int dst, bool second_is_vel, bool has_second, bool with_forces) {
  MAVI_LAUNCH(c, k_gather, nblk(p.n), TPB, p, perm, cell_new, start, a.pos[src], a.pos[dst],
              (has_second && second_is_vel) ? a.vel[src] : nullptr, (has_second && second_is_vel) ? a.vel[dst] : nullptr,
              (has_second && !second_is_vel) ? a.ang[src] : nullptr, (has_second && !second_is_vel) ? a.ang[dst] : nullptr,
              a.idflag[src], a.idflag[dst], a.cell[dst], with_forces ? a.force : nullptr,
              with_forces ? a.force_old : nullptr);
}

// =========================================================================================================
// Pair forces (calc_forces!, src/integration.jl:112-224) as a per-particle gather, fused with the integrators.
// =========================================================================================================

// flags[] layout (device error / control word)
enum { FLAG_ERR = 0, FLAG_CHANGED = 1, FLAG_BIGMOVE = 2, FLAG_NFIX = 3 };

template <int DYN, bool MINIMG>
__device__ __forceinline__ void accumulate_pair(const DevParams &p, double2 ri, double2 rj, double &fx, double &fy) {
  double dx = min_image<MINIMG>(ri.x - rj.x, p.half[0], p.size[0]);
  double dy = min_image<MINIMG>(ri.y - rj.y, p.half[1], p.size[1]);
  // laws with a cutoff need r2 rounded exactly like the reference (bit-exact neighbour decisions); LJ has no cutoff
  double r2 = (DYN == MAVI_DYN_LJ) ? fma(dx, dx, dy * dy) : dist2_exact(dx, dy);
  double c = pair_coef<DYN>(p, r2);
  fx = fma(c, dx, fx);
  fy = fma(c, dy, fy);
}

// Interior cell (no stencil wrap / clipping): the neighbours are 4 contiguous runs of the sorted order
//   column-1: [a0,b0)   own column: [a1,k) and (k,b1)   column+1: [a2,b2)
// walked by ONE flat, branch-free loop: neighbour t lives at slot t + offset(t) (three compares select the offset), so
// a warp waits for the max over lanes of the neighbour COUNT instead of the sum of per-column maxima, and the next
// position is prefetched while the current pair is evaluated.
template <int DYN, bool MINIMG>
__device__ __forceinline__ void interior_force(const DevParams &p, const int *__restrict__ start,
                                               const double2 *__restrict__ pos, int cell, int k, double2 ri, double &fx,
                                               double &fy) {
  const int R = p.num_rows;
  const int *s = start + cell;
  const int a0 = __ldg(s - R - 1), b0 = __ldg(s - R + 2);
  const int a1 = __ldg(s - 1), b1 = __ldg(s + 2);
  const int a2 = __ldg(s + R - 1), b2 = __ldg(s + R + 2);
  const int c1 = b0 - a0;            // neighbours t <  c1          -> slot a0 + t
  const int c2 = c1 + (k - a1);      //            c1 <= t < c2    -> slot a1 + (t - c1)
  const int c3 = c2 + (b1 - k - 1);  //            c2 <= t < c3    -> slot k + 1 + (t - c2)   (skips self)
  const int total = c3 + (b2 - a2);  //            c3 <= t         -> slot a2 + (t - c3)
  if (total <= 0) return;
  const int d1 = (a1 - c1) - a0, d3 = (a2 - c3) - (a1 - c1) - 1;
  // slot(t) = t + a0 + [t>=c1] d1 + [t>=c2] + [t>=c3] d3   (predicated adds, no branches)
  auto slot = [&](int t) { return t + a0 + (t >= c1 ? d1 : 0) + (t >= c2 ? 1 : 0) + (t >= c3 ? d3 : 0); };
  // two neighbours per trip: independent FP64 chains, and the loads of the next pair are in flight meanwhile
  double2 r0 = __ldg(pos + slot(0));
  double2 r1 = (total > 1) ? __ldg(pos + slot(1)) : r0;
  int t = 0;
#pragma unroll 1
  for (; t + 2 <= total; t += 2) {
    const double2 q0 = r0, q1 = r1;
    if (t + 2 < total) r0 = __ldg(pos + slot(t + 2));
    if (t + 3 < total) r1 = __ldg(pos + slot(t + 3));
    accumulate_pair<DYN, MINIMG>(p, ri, q0, fx, fy);
    accumulate_pair<DYN, MINIMG>(p, ri, q1, fx, fy);
  }
  if (t < total) accumulate_pair<DYN, MINIMG>(p, ri, r0, fx, fy);
}

// ALLP: chunks === nothing -> all pairs over active ids (src/integration.jl:197-224); physical order = id order.
// exact_minimg: force the minimum image on interior cells too (pass B after an abnormally large drift).
template <int DYN, bool PER, bool ALLP>
__device__ __forceinline__ double2 pair_force(const DevParams &p, const int *__restrict__ start,
                                              const double2 *__restrict__ pos, const unsigned int *__restrict__ idflag,
                                              int cell, int k, double2 ri, bool exact_minimg) {
  double fx = 0.0, fy = 0.0;
  if (ALLP) {
    for (int j = 0; j < p.n; j++) {
      if (j == k || (idflag[j] & MAVI_INACTIVE_BIT)) continue;
      accumulate_pair<DYN, PER>(p, ri, __ldg(pos + j), fx, fy);
    }
  } else {
    const int R = p.num_rows;
    const int col = cell / R, row = cell - col * R;
    const bool interior = row >= 1 && row <= R - 2 && col >= 1 && col <= p.num_cols - 2;
    if (interior) {
      // Fresh cells: both particles of a pair lie inside 8-adjacent cells, so |dr| < 2 cell widths <= size/4 on a grid
      // of >= 8 cells per axis and the reference's `abs(dr) > size/2` test is false: min image skipped EXACTLY.
      // Stale cells (Verlet pass 2) are covered by the per-step displacement guard (FLAG_BIGMOVE).
      if (PER && (exact_minimg || !p.fast_interior)) interior_force<DYN, true>(p, start, pos, cell, k, ri, fx, fy);
      else interior_force<DYN, false>(p, start, pos, cell, k, ri, fx, fy);
    } else {
      for_each_neighbor(p, start, cell, k, [&](int j) { accumulate_pair<DYN, PER>(p, ri, __ldg(pos + j), fx, fy); });
    }
  }
  return make_double2(fx, fy);
}

// clean_forces! + calc_forces! (+ calc_walls_forces!): the force state after src/integration.jl:508-511.
template <int DYN, bool PER, bool ALLP>
__global__ void __launch_bounds__(TPB) k_force_only(const __grid_constant__ DevParams p, const int *__restrict__ start,
                             const int *__restrict__ cell, const unsigned int *__restrict__ idflag,
                             const double2 *__restrict__ pos, double2 *__restrict__ force, int with_walls) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n) return;
  double2 F = make_double2(0.0, 0.0);
  if (!(idflag[k] & MAVI_INACTIVE_BIT)) {
    double2 r = pos[k];
    F = pair_force<DYN, PER, ALLP>(p, start, pos, idflag, ALLP ? 0 : cell[k], k, r, false);
    if (with_walls && p.has_force_walls) wall_forces(p, r.x, r.y, F.x, F.y);
  }
  force[k] = F;
}

// newton_step! first half (src/integration.jl:507-512 + update_verlet! :418-424):
//   F1 = pair forces + wall forces;  pos' = pos + vel dt + F1 dt^2/2  (every slot, active or not).
template <int DYN, bool PER, bool ALLP>
__global__ void __launch_bounds__(TPB) k_newton_a(const __grid_constant__ DevParams p, const int *__restrict__ start,
                           const int *__restrict__ cell, const unsigned int *__restrict__ idflag,
                           const double2 *__restrict__ pos_in, const double2 *__restrict__ vel,
                           double2 *__restrict__ pos_out, double2 *__restrict__ f1, int *__restrict__ flags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n) return;
  double2 r = pos_in[k];
  double2 F = make_double2(0.0, 0.0);
  if (!(idflag[k] & MAVI_INACTIVE_BIT)) {
    F = pair_force<DYN, PER, ALLP>(p, start, pos_in, idflag, ALLP ? 0 : cell[k], k, r, false);
    if (p.has_force_walls) wall_forces(p, r.x, r.y, F.x, F.y);
  }
  double2 v = vel[k];
  double mx = v.x * p.dt + F.x * p.term, my = v.y * p.dt + F.y * p.term;
  r.x = r.x + mx;
  r.y = r.y + my;
  // displacement guard of the interior fast path: a drift beyond one cell in one step makes pass B take the exact
  // (minimum-image everywhere) path.  Never happens in a stable run; keeps the fast path assumption-free.
  if (!ALLP && PER && !(fabs(mx) <= p.cl && fabs(my) <= p.ch)) flags[FLAG_BIGMOVE] = 1;
  pos_out[k] = r;
  f1[k] = F;
}

// newton_step! second half (update_verlet! :426-430, walls! :513): F2 on the drifted positions with the STALE cell
// lists and WITHOUT wall forces; vel += dt/2 (F2 + F1); walls!(active ids).  Also decides, exactly, whether the next
// update_chunks! would put the particle into the cell it is sorted under (-> the re-sort can be skipped), and defers
// position changes made by walls! (periodic wrap, slippery projection) to a sparse fix-up list because neighbours
// still read the unmodified drifted positions in this launch.
template <int DYN, bool PER, bool ALLP>
__global__ void __launch_bounds__(TPB) k_newton_b(const __grid_constant__ DevParams p, const int *__restrict__ start,
                           const int *__restrict__ cell, const unsigned int *__restrict__ idflag,
                           const double2 *__restrict__ pos_in, double2 *__restrict__ vel,
                           const double2 *__restrict__ f1, double2 *__restrict__ f2, int *__restrict__ flags,
                           int *__restrict__ fix_idx, double2 *__restrict__ fix_pos) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  bool changed = false;
  if (k < p.n) {
    double2 r = pos_in[k];
    double2 F = make_double2(0.0, 0.0);
    const bool active = !(idflag[k] & MAVI_INACTIVE_BIT);
    const int c = ALLP ? 0 : cell[k];
    if (active) F = pair_force<DYN, PER, ALLP>(p, start, pos_in, idflag, c, k, r, !ALLP && flags[FLAG_BIGMOVE] != 0);
    double2 v = vel[k];
    double2 Fo = f1[k];
    v.x = v.x + p.hdt * (F.x + Fo.x);
    v.y = v.y + p.hdt * (F.y + Fo.y);
    if (active) {
      const double x0 = r.x, y0 = r.y;
      apply_walls<true>(p, r.x, r.y, v.x, v.y, p.particle_radius);
      if (r.x != x0 || r.y != y0) {
        int m = atomicAdd(&flags[FLAG_NFIX], 1);
        fix_idx[m] = k;
        fix_pos[m] = r;
      }
      if (!ALLP) changed = !still_in_cell(p, r.x, r.y, c);
    }
    vel[k] = v;
    f2[k] = F;
  }
  unsigned int m = __ballot_sync(0xffffffffu, changed);
  if (m && (threadIdx.x & 31) == 0) atomicAdd(&flags[FLAG_CHANGED], __popc(m));
}

__global__ void k_apply_pos_fixes(const int *__restrict__ flags, const int *__restrict__ fix_idx,
                                  const double2 *__restrict__ fix_pos, double2 *__restrict__ pos) {
  const int n = flags[FLAG_NFIX];
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) pos[fix_idx[m]] = fix_pos[m];
}

// szabo_step! / rtp_step! (src/integration.jl:517-535): forces + update_szabo! (:433-465) / update_rtp! (:467-498)
// + walls! in ONE pass.  The update loops slots 1:count (not ids) like the reference.
template <int DYN, bool PER, bool ALLP>
__global__ void __launch_bounds__(TPB) k_self_propelled(const __grid_constant__ DevParams p, const int *__restrict__ start,
                                 const int *__restrict__ cell, const unsigned int *__restrict__ idflag,
                                 const double2 *__restrict__ pos_in, double *__restrict__ ang,
                                 double2 *__restrict__ pos_out, double2 *__restrict__ force,
                                 const double *__restrict__ noise, unsigned long long step, int *__restrict__ flags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  bool changed = false;
  if (k < p.n) {
    const unsigned int idf = idflag[k];
    const bool active = !(idf & MAVI_INACTIVE_BIT);
    const unsigned int id = idf & ~MAVI_INACTIVE_BIT;
    const int c = ALLP ? 0 : cell[k];
    double2 r = pos_in[k];
    double2 F = make_double2(0.0, 0.0);
    if (active) {
      F = pair_force<DYN, PER, ALLP>(p, start, pos_in, idflag, c, k, r, false);
      if (p.has_force_walls) wall_forces(p, r.x, r.y, F.x, F.y);
    }
    force[k] = F;
    if ((int)id < p.n_count) {
      double theta = ang[k];
      double sn, cs;
      sincos(theta, &sn, &cs);
      if (DYN == MAVI_DYN_SZABO) {
        const double vo = p.dyn[0], mu = p.dyn[1], relax_time = p.dyn[2], drot = p.dyn[7];
        double velx = vo * cs + mu * F.x, vely = vo * sn + mu * F.y;
        double speed = sqrt(fabs(velx) + fabs(vely));  // sqrt(sum(abs, vel)) (sic), :448
        double cross_prod = speed > 0.0 ? (cs * vely - sn * velx) / speed : 0.0;
        if (fabs(cross_prod) > 1.0) cross_prod = sign_d(cross_prod);
        double nz = 0.0;
        if (drot != 0.0) nz = (p.rng_mode == MAVI_RNG_HOST_NOISE) ? (noise ? noise[id] : 0.0) : philox_normal(p.seed, id, step);
        double d_theta = 1.0 / relax_time * asin(cross_prod) * p.dt + sqrt(2.0 * drot * p.dt) * nz;
        r.x += velx * p.dt;
        r.y += vely * p.dt;
        ang[k] = theta + d_theta;
      } else {
        const double vo = p.dyn[0], tumble_rate = p.dyn[3];
        double velx = vo * cs + F.x, vely = vo * sn + F.y;
        r.x += velx * p.dt;
        r.y += vely * p.dt;
        double u, u2;
        if (p.rng_mode == MAVI_RNG_HOST_NOISE) {
          u = noise ? noise[2 * (size_t)id] : 1.0;
          u2 = noise ? noise[2 * (size_t)id + 1] : 0.0;
        } else {
          philox_uniform2(p.seed, id, step, u, u2);
        }
        if (u < tumble_rate * p.dt) ang[k] = 6.283185307179586 * u2;  // 2*pi*rand(), :495
      }
    }
    if (active) {
      double vx = 0.0, vy = 0.0;
      apply_walls<false>(p, r.x, r.y, vx, vy, p.particle_radius);
      if (!ALLP) changed = !still_in_cell(p, r.x, r.y, c);
    }
    pos_out[k] = r;
  }
  unsigned int m = __ballot_sync(0xffffffffu, changed);
  if (m && (threadIdx.x & 31) == 0) atomicAdd(&flags[FLAG_CHANGED], __popc(m));
}

// ---- dispatch over (dynamics, periodic, all-pairs) ----------------------------------------------------------
#define MAVI_DISPATCH_DYN(DYNV, PERV, ALLPV, CALL)                                    \
  do {                                                                                \
    if (PERV) {                                                                       \
      if (ALLPV) { CALL(DYNV, true, true); } else { CALL(DYNV, true, false); }        \
    } else {                                                                          \
      if (ALLPV) { CALL(DYNV, false, true); } else { CALL(DYNV, false, false); }      \
    }                                                                                 \
  } while (0)

void launch_force_only(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int cur, bool with_wall_forces) {
  const bool allp = p.num_cells == 0;
#define CALL(D, P, A) \
  MAVI_LAUNCH(c, (k_force_only<D, P, A>), nblk(p.n), TPB, p, a.start, a.cell[cur], a.idflag[cur], a.pos[cur], a.force, (int)with_wall_forces)
  switch (p.dynamics) {
    case MAVI_DYN_LJ: MAVI_DISPATCH_DYN(MAVI_DYN_LJ, p.periodic, allp, CALL); break;
    case MAVI_DYN_HARMTRUNC: MAVI_DISPATCH_DYN(MAVI_DYN_HARMTRUNC, p.periodic, allp, CALL); break;
    case MAVI_DYN_SZABO: MAVI_DISPATCH_DYN(MAVI_DYN_SZABO, p.periodic, allp, CALL); break;
    case MAVI_DYN_RTP: MAVI_DISPATCH_DYN(MAVI_DYN_RTP, p.periodic, allp, CALL); break;
  }
#undef CALL
}

void launch_newton_a(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int cur) {
  const bool allp = p.num_cells == 0;
#define CALL(D, P, A) \
  MAVI_LAUNCH(c, (k_newton_a<D, P, A>), nblk(p.n), TPB, p, a.start, a.cell[cur], a.idflag[cur], a.pos[cur], a.vel[cur], a.pos[cur ^ 1], a.force_old, a.flags)
  if (p.dynamics == MAVI_DYN_LJ) MAVI_DISPATCH_DYN(MAVI_DYN_LJ, p.periodic, allp, CALL);
  else MAVI_DISPATCH_DYN(MAVI_DYN_HARMTRUNC, p.periodic, allp, CALL);
#undef CALL
}

void launch_newton_b(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int cur) {
  const bool allp = p.num_cells == 0;
  // reads the drifted positions pos[cur^1] (which become the current positions after the sparse wall fix-ups)
#define CALL(D, P, A) \
  MAVI_LAUNCH(c, (k_newton_b<D, P, A>), nblk(p.n), TPB, p, a.start, a.cell[cur], a.idflag[cur], a.pos[cur ^ 1], a.vel[cur], a.force_old, a.force, a.flags, a.fix_idx, a.fix_pos)
  if (p.dynamics == MAVI_DYN_LJ) MAVI_DISPATCH_DYN(MAVI_DYN_LJ, p.periodic, allp, CALL);
  else MAVI_DISPATCH_DYN(MAVI_DYN_HARMTRUNC, p.periodic, allp, CALL);
#undef CALL
  MAVI_LAUNCH(c, k_apply_pos_fixes, 64, TPB, a.flags, a.fix_idx, a.fix_pos, a.pos[cur ^ 1]);
}

void launch_self_propelled(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int cur, const double *noise,
                           unsigned long long step) {
  const bool allp = p.num_cells == 0;
#define CALL(D, P, A) \
  MAVI_LAUNCH(c, (k_self_propelled<D, P, A>), nblk(p.n), TPB, p, a.start, a.cell[cur], a.idflag[cur], a.pos[cur], a.ang[cur], a.pos[cur ^ 1], a.force, noise, step, a.flags)
  if (p.dynamics == MAVI_DYN_SZABO) MAVI_DISPATCH_DYN(MAVI_DYN_SZABO, p.periodic, allp, CALL);
  else MAVI_DISPATCH_DYN(MAVI_DYN_RTP, p.periodic, allp, CALL);
#undef CALL
}

// =========================================================================================================
// Quantities (src/quantities.jl) as deterministic two-stage block reductions.
// =========================================================================================================
constexpr int RED_TPB = 256, RED_MAX_BLOCKS = 1184;  // 8 CTAs x 148 SMs

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double ws[RED_TPB / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < RED_TPB / 32) t = ws[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in thread 0
}

__global__ void k_reduce_final(const double *__restrict__ partials, int nb, double scale, double *__restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partials[i];
  s = block_sum(s);
  if (threadIdx.x == 0) *out = s * scale;
}

// kinetic_energy, src/quantities.jl:12-18: sum over ALL slots of |v|^2, /2 (mass 1)
__global__ void k_kinetic(int n, const double2 *__restrict__ vel, double *__restrict__ partials) {
  double s = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    double2 v = vel[k];
    s += v.x * v.x + v.y * v.y;
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

void launch_kinetic_energy(const LaunchCtx &c, const DevParams &p, const double2 *vel, double *partials, double *out) {
  int nb = min(nblk(p.n, RED_TPB), RED_MAX_BLOCKS);
  if (nb < 1) nb = 1;
  MAVI_LAUNCH(c, k_kinetic, nb, RED_TPB, p.n, vel, partials);
  MAVI_LAUNCH(c, k_reduce_final, 1, RED_TPB, partials, nb, 0.5, out);
}

// potential_energy(::LenJonesCfg), src/quantities.jl:46-66.  MODE 0: every pair i<j of slots 1:count (exact, O(N^2));
// MODE 1: the cell-stencil pair set (each pair seen from both ends -> halved).
template <bool PER, int MODE>
__global__ void k_potential(const __grid_constant__ DevParams p, const int *__restrict__ start,
                            const int *__restrict__ cell, const unsigned int *__restrict__ idflag,
                            const double2 *__restrict__ pos, double *__restrict__ partials) {
  double s = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < p.n; k += gridDim.x * blockDim.x) {
    const unsigned int idf = idflag[k];
    double2 ri = pos[k];
    auto term = [&](int j) {
      double2 rj = __ldg(pos + j);
      double dx = min_image<PER>(ri.x - rj.x, p.half[0], p.size[0]);
      double dy = min_image<PER>(ri.y - rj.y, p.half[1], p.size[1]);
      double s2 = p.lj_sig2 / (dx * dx + dy * dy);
      double s6 = s2 * s2 * s2;
      s += s6 * s6 - s6;
    };
    if (MODE == 0) {
      if ((int)(idf & ~MAVI_INACTIVE_BIT) >= p.n_count) continue;
      for (int j = k + 1; j < p.n; j++)
        if ((int)(idflag[j] & ~MAVI_INACTIVE_BIT) < p.n_count) term(j);
    } else {
      if (idf & MAVI_INACTIVE_BIT) continue;
      for_each_neighbor(p, start, cell[k], k, term);
    }
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

void launch_potential_energy(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int cur, int mode, double *out) {
  int nb = min(nblk(p.n, RED_TPB), RED_MAX_BLOCKS);
  if (nb < 1) nb = 1;
  const double eps4 = 4.0 * p.dyn[1];
  if (mode == 0) {
    if (p.periodic) MAVI_LAUNCH(c, (k_potential<true, 0>), nb, RED_TPB, p, a.start, a.cell[cur], a.idflag[cur], a.pos[cur], a.reduce_buf);
    else MAVI_LAUNCH(c, (k_potential<false, 0>), nb, RED_TPB, p, a.start, a.cell[cur], a.idflag[cur], a.pos[cur], a.reduce_buf);
    MAVI_LAUNCH(c, k_reduce_final, 1, RED_TPB, a.reduce_buf, nb, eps4, out);
  } else {
    if (p.periodic) MAVI_LAUNCH(c, (k_potential<true, 1>), nb, RED_TPB, p, a.start, a.cell[cur], a.idflag[cur], a.pos[cur], a.reduce_buf);
    else MAVI_LAUNCH(c, (k_potential<false, 1>), nb, RED_TPB, p, a.start, a.cell[cur], a.idflag[cur], a.pos[cur], a.reduce_buf);
    MAVI_LAUNCH(c, k_reduce_final, 1, RED_TPB, a.reduce_buf, nb, 0.5 * eps4, out);
  }
}

// =========================================================================================================
// Upload / download helpers
// =========================================================================================================
__global__ void k_unpermute2(int n, const unsigned int *__restrict__ idflag, const double2 *__restrict__ in,
                             double2 *__restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[idflag[k] & ~MAVI_INACTIVE_BIT] = in[k];
}
__global__ void k_unpermute1(int n, const unsigned int *__restrict__ idflag, const double *__restrict__ in,
                             double *__restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[idflag[k] & ~MAVI_INACTIVE_BIT] = in[k];
}
__global__ void k_unpermute_cells(int n, int num_cells, const unsigned int *__restrict__ idflag,
                                  const int *__restrict__ cell, int *__restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    int c = cell[k];
    out[idflag[k] & ~MAVI_INACTIVE_BIT] = c < num_cells ? c : -1;
  }
}
__global__ void k_ids(int n, const unsigned int *__restrict__ idflag, int *__restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = (int)(idflag[k] & ~MAVI_INACTIVE_BIT);
}
__global__ void k_init_ids(int n, const unsigned char *__restrict__ mask, unsigned int *__restrict__ idflag,
                           int *__restrict__ cell, int num_cells) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    bool act = mask ? mask[k] != 0 : true;
    idflag[k] = (unsigned int)k | (act ? 0u : MAVI_INACTIVE_BIT);
    cell[k] = -2;  // "not sorted yet": forces the first re-sort
  }
}
// check_inside, src/space_checks.jl:9-61 (Rectangle: any coordinate < bottom_left or > top_right; Circle: |pos|^2 > R^2,
// centre ignored (sic)); only single-geometry spaces are checked (ManyGeometries hits the generic no-op method).
__global__ void k_check_inside(const __grid_constant__ DevParams p, const double2 *__restrict__ pos,
                               const unsigned int *__restrict__ idflag, int *__restrict__ flags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n || (idflag[k] & MAVI_INACTIVE_BIT)) return;
  const DevSpace &sp = p.spaces[0];
  double2 r = pos[k];
  bool out = false;
  if (sp.geom == MAVI_GEOM_RECT) {
    double trx = sp.rect_bl[0] + sp.rect_sz[0], try_ = sp.rect_bl[1] + sp.rect_sz[1];
    out = (r.x < sp.rect_bl[0]) || (r.y < sp.rect_bl[1]) || (r.x > trx) || (r.y > try_);
  } else if (sp.geom == MAVI_GEOM_CIRCLE) {
    out = (r.x * r.x + r.y * r.y) > sp.cr * sp.cr;
  }
  if (out) atomicOr(&flags[0], ERRBIT_OUTSIDE_SPACE);
}

void launch_unpermute2(const LaunchCtx &c, int n, const unsigned int *idflag, const double2 *in, double2 *out) {
  MAVI_LAUNCH(c, k_unpermute2, nblk(n), TPB, n, idflag, in, out);
}
void launch_unpermute1(const LaunchCtx &c, int n, const unsigned int *idflag, const double *in, double *out) {
  MAVI_LAUNCH(c, k_unpermute1, nblk(n), TPB, n, idflag, in, out);
}
void launch_unpermute_cells(const LaunchCtx &c, int n, int num_cells, const unsigned int *idflag, const int *cell, int *out) {
  MAVI_LAUNCH(c, k_unpermute_cells, nblk(n), TPB, n, num_cells, idflag, cell, out);
}
void launch_ids(const LaunchCtx &c, int n, const unsigned int *idflag, int *out) {
  MAVI_LAUNCH(c, k_ids, nblk(n), TPB, n, idflag, out);
}
void launch_init_ids(const LaunchCtx &c, int n, const unsigned char *mask, unsigned int *idflag, int *cell, int num_cells) {
  MAVI_LAUNCH(c, k_init_ids, nblk(n), TPB, n, mask, idflag, cell, num_cells);
}
void launch_check_inside(const LaunchCtx &c, const DevParams &p, const double2 *pos, const unsigned int *idflag, int *flags) {
  MAVI_LAUNCH(c, k_check_inside, nblk(p.n), TPB, p, pos, idflag, flags);
}

}  // namespace mavi
