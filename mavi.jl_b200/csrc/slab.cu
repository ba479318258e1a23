// slab.cu — x-slab domain decomposition over the GPUs of one node: one process per GPU, NCCL send/recv over
// NVLink 5 / NVSwitch for the two exchange steps the path really has:
//   halo     : the boundary cell column of each neighbour (positions + its tstart rows), before EVERY force pass
//   migration: particles whose cell column left the slab, once per step, merged by the same incremental tile repair
// The reference has no distributed path; it partitions the same loop over cell columns across threads
// (src/integration.jl:159-194).  Global cell ids stay those of the single-GPU / reference binning.
//
// Local grid = [left halo column | m owned columns | right halo column]; a cell column is ONE contiguous slab of
// tpc*cap slots, so a halo or an emigrant column is a single contiguous message per array.
#include <dlfcn.h>

#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "handle.cuh"

namespace mavi {

// ---- minimal NCCL surface, resolved at run time so that single-GPU users need no NCCL at all ---------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclUint8 = 1 };
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static const char *load_nccl() {
  if (g_nccl.lib) return nullptr;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return "libnccl.so.2 not found";
#define NCCL_SYM(field, name)                                   \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);          \
  if (!g_nccl.field) return "missing NCCL symbol " name;
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  NCCL_SYM(CommInitRank, "ncclCommInitRank")
  NCCL_SYM(CommDestroy, "ncclCommDestroy")
  NCCL_SYM(GroupStart, "ncclGroupStart")
  NCCL_SYM(GroupEnd, "ncclGroupEnd")
  NCCL_SYM(Send, "ncclSend")
  NCCL_SYM(Recv, "ncclRecv")
  NCCL_SYM(AllReduce, "ncclAllReduce")
  NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
  return nullptr;
}

#define SLAB_CUDA(h, expr)                                                                              \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      (h)->set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #expr); \
      return MAVI_ERR_CUDA;                                                                             \
    }                                                                                                   \
  } while (0)
#define SLAB_NCCL(h, expr)                                                                                 \
  do {                                                                                                     \
    ncclResult_t r_ = (expr);                                                                              \
    if (r_ != 0) {                                                                                         \
      (h)->set_error("NCCL error %s at %s:%d (%s)", g_nccl.GetErrorString(r_), __FILE__, __LINE__, #expr); \
      return MAVI_ERR_NCCL;                                                                                \
    }                                                                                                      \
  } while (0)

int slab_unique_id(void *out128) {
  const char *e = load_nccl();
  if (e) return MAVI_ERR_NCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return MAVI_ERR_NCCL;
  memcpy(out128, &id, sizeof id);
  return MAVI_OK;
}

// columns owned by `rank`: contiguous, the first C % world ranks get one extra
static void slab_columns(int C, int world, int rank, int *lo, int *m) {
  const int base = C / world, rem = C % world;
  *m = base + (rank < rem ? 1 : 0);
  *lo = rank * base + (rank < rem ? rank : rem);
}

// called from validate_and_lower once the global chunk parameters are known
int slab_configure(Handle *h, const MaviParams *mp) {
  DevParams &p = h->p;
  SlabState &s = h->slab;
  if (mp->n_spaces != 1 || mp->spaces[0].wall != MAVI_WALL_PERIODIC || mp->spaces[0].geom != MAVI_GEOM_RECT ||
      mp->num_cols <= 0) {
    h->set_error("multi-GPU slabs need a single periodic rectangle with chunks (SURVEY.md 8e: all-pairs mode and "
                 "composite spaces are 'replicas only')");
    return MAVI_ERR_UNSUPPORTED;
  }
  if (mp->dynamics == MAVI_DYN_RINGS) {
    h->set_error("Mavi.Rings runs on one GPU in this version (a ring may straddle a slab boundary)");
    return MAVI_ERR_UNSUPPORTED;
  }
  if (!mp->nccl_unique_id) {
    h->set_error("world > 1 needs MaviParams.nccl_unique_id (mavi_nccl_unique_id on rank 0, broadcast by the host)");
    return MAVI_ERR_BAD_PARAMS;
  }
  const char *e = load_nccl();
  if (e) {
    h->set_error("NCCL unavailable: %s", e);
    return MAVI_ERR_NCCL;
  }
  s.world = mp->world;
  s.n_global = (int)(mp->n_global > 0 ? mp->n_global : mp->n);
  s.rank = mp->rank;
  const int C = mp->num_cols;
  slab_columns(C, s.world, s.rank, &s.col_lo, &s.m);
  if (s.m < 2) {
    h->set_error("every slab needs at least 2 cell columns (%d columns over %d GPUs)", C, s.world);
    return MAVI_ERR_BAD_PARAMS;
  }
  s.left = (s.rank + s.world - 1) % s.world;
  s.right = (s.rank + 1) % s.world;
  int lo;
  slab_columns(C, s.world, s.left, &lo, &s.m_left);
  slab_columns(C, s.world, s.right, &lo, &s.m_right);
  p.slab = 1;
  p.gcols = C;
  p.col_lo = s.col_lo;
  p.num_cols = s.m + 2;
  p.num_cells = p.num_cols * p.num_rows;
  p.ord_cols = s.m;
  p.ord_col0 = 1;
  p.wrap_cols = 0;
  p.wrap_rows = 1;
  p.seam_left = s.rank == 0;
  p.seam_right = s.rank == s.world - 1;
  ncclUniqueId id;
  memcpy(&id, mp->nccl_unique_id, sizeof id);
  SLAB_NCCL(h, g_nccl.CommInitRank(&s.comm, s.world, id, s.rank));
  return MAVI_OK;
}

// max over all ranks of one int (tile capacity agreement: every rank must use the SAME slots-per-column, because a
// halo / emigrant column is shipped as one raw slab of tpc*cap slots)
int slab_allreduce_max(Handle *h, int *value) {
  int *d = h->a.flags + FLAG_SCRATCH;
  SLAB_CUDA(h, cudaMemcpyAsync(d, value, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  SLAB_NCCL(h, g_nccl.AllReduce(d, d, 1, /*ncclInt32*/ 2, /*ncclMax*/ 2, h->slab.comm, h->stream));
  SLAB_CUDA(h, cudaMemcpyAsync(value, d, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  SLAB_CUDA(h, cudaStreamSynchronize(h->stream));
  return MAVI_OK;
}

void slab_destroy(Handle *h) {
  if (h->slab.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->slab.comm);
  h->slab.comm = nullptr;
}

// ---- kernels ------------------------------------------------------------------------------------------------------
constexpr int TR1 = MAVI_TR + 1;

// halo tstart rows arrive as the SENDER's absolute slots: shift them by (receiver column base - sender column base)
__global__ void k_rebase_tstart(int *__restrict__ ts, int n, int delta) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ts[i] += delta;
}

// empty the tiles of one local column (tstart rows -> tile base)
__global__ void k_clear_column(const __grid_constant__ DevParams p, int *__restrict__ tstart, int lcol) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.tpc * TR1) return;
  const int tr = i / TR1;
  tstart[(size_t)lcol * p.tpc * TR1 + i] = (lcol * p.tpc + tr) * p.cap;
}

// Immigrants: every received record is queued as an inter-tile mover of THIS rank: record -> mover list, destination
// tile inbox, tile marked dirty.  The incremental repair then merges it exactly like a local mover.
__global__ void k_ingest(const __grid_constant__ DevParams p, const EmRec *__restrict__ recs, int count_flag, int em_cap,
                         int has_vel, double2 *__restrict__ mv_pos, double2 *__restrict__ mv_second,
                         double2 *__restrict__ mv_force, unsigned int *__restrict__ mv_id, int *__restrict__ mv_cell,
                         int *__restrict__ mv_src, int *__restrict__ tile_dirty, int *__restrict__ dirty_list,
                         int *__restrict__ inbox_cnt, int *__restrict__ inbox, int *__restrict__ flags,
                         int *__restrict__ chg) {
  if (!flags[FLAG_RAN]) return;
  const int n = min(flags[count_flag], em_cap);
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n; l += gridDim.x * blockDim.x) {
    const EmRec e = recs[l];
    int c = cell_of_point(p, e.pos.x, e.pos.y);
    const int lcol = c >= 0 ? div_rows(p, c) : -1;
    if (c < 0 || lcol < 1 || lcol > p.num_cols - 2) {
      atomicOr(&flags[FLAG_ERR], ERRBIT_OUT_OF_GRID);
      continue;
    }
    const int t = tile_of_cell(p, c);
    if (atomicExch(&tile_dirty[t], 1) == 0) dirty_list[atomicAdd(&flags[FLAG_CHANGED], 1)] = t;
    if (chg) {  // force carry: the arrival changes the membership of its cell
      const int q = atomicAdd(&flags[FLAG_NCHG], 2);
      if (q + 1 < p.chg_cap) { chg[q] = c; chg[q + 1] = c; }
      else atomicOr(&flags[FLAG_OVERFLOW], 8);
    }
    const int m = atomicAdd(&flags[FLAG_NMV], 1);
    const int i = atomicAdd(&inbox_cnt[t], 1);
    if (m < p.mv_cap && i < p.inbox_cap) {
      mv_src[m] = -1;  // record already in place (k_repair_collect skips it)
      mv_pos[m] = e.pos;
      mv_second[m] = e.second;
      mv_force[m] = e.force;
      mv_id[m] = e.idflag;
      mv_cell[m] = c;
      inbox[(size_t)t * p.inbox_cap + i] = m;
    } else {
      atomicOr(&flags[FLAG_OVERFLOW], m < p.mv_cap ? 2 : 4);
    }
  }
}

__global__ void k_owned_count(const __grid_constant__ DevParams p, const int *__restrict__ tile_prefix, int *__restrict__ flags) {
  flags[FLAG_TAIL] = tile_prefix[p.nt_ord];
}

// ---- exchanges ------------------------------------------------------------------------------------------------------
static size_t col_slots(const DevParams &p) { return (size_t)p.tpc * p.cap; }

// halo positions (pos_buf = a.pos[0] or a.pos[1]); with_layout also ships the tstart rows of the boundary columns
int slab_halo_exchange(Handle *h, double2 *pos_buf, bool with_layout) {
  const DevParams &p = h->p;
  SlabState &s = h->slab;
  DevArrays &a = h->a;
  const size_t cs = col_slots(p), m = (size_t)s.m;
  const size_t rows = (size_t)p.tpc * TR1;
  // my boundary columns: local 1 (-> left neighbour's right halo) and local m (-> right neighbour's left halo)
  SLAB_NCCL(h, g_nccl.GroupStart());
  SLAB_NCCL(h, g_nccl.Send(pos_buf + 1 * cs, cs * sizeof(double2), ncclInt8, s.left, s.comm, h->stream));
  SLAB_NCCL(h, g_nccl.Send(pos_buf + m * cs, cs * sizeof(double2), ncclInt8, s.right, s.comm, h->stream));
  // receive order matters when left == right (2 GPUs): the peer's FIRST send is its column 1 = my RIGHT halo
  SLAB_NCCL(h, g_nccl.Recv(pos_buf + (m + 1) * cs, cs * sizeof(double2), ncclInt8, s.right, s.comm, h->stream));
  SLAB_NCCL(h, g_nccl.Recv(pos_buf + 0 * cs, cs * sizeof(double2), ncclInt8, s.left, s.comm, h->stream));
  if (with_layout) {
    SLAB_NCCL(h, g_nccl.Send(a.tstart + 1 * rows, rows * sizeof(int), ncclInt8, s.left, s.comm, h->stream));
    SLAB_NCCL(h, g_nccl.Send(a.tstart + m * rows, rows * sizeof(int), ncclInt8, s.right, s.comm, h->stream));
    SLAB_NCCL(h, g_nccl.Recv(a.tstart + (m + 1) * rows, rows * sizeof(int), ncclInt8, s.right, s.comm, h->stream));
    SLAB_NCCL(h, g_nccl.Recv(a.tstart + 0 * rows, rows * sizeof(int), ncclInt8, s.left, s.comm, h->stream));
  }
  SLAB_NCCL(h, g_nccl.GroupEnd());
  if (with_layout) {
    // right halo came from the right neighbour's column 1; left halo from the left neighbour's column m_left
    const int nb = ((int)rows + 255) / 256;
    k_rebase_tstart<<<nb, 256, 0, h->stream>>>(a.tstart + (m + 1) * rows, (int)rows, (int)((m + 1) * cs) - (int)(1 * cs));
    k_rebase_tstart<<<nb, 256, 0, h->stream>>>(a.tstart + 0 * rows, (int)rows, 0 - (int)((size_t)s.m_left * cs));
    h->launches += 2;
  }
  return MAVI_OK;
}

// Migration: the integrate kernel wrote the record of every particle that crossed into a halo column straight into
// the emigrant buffer of that side (note_if_moved).  Ship both buffers (fixed capacity, the count travels separately)
// and queue what arrived as movers; the ONE tile repair of the step then handles local movers and immigrants alike.
static int slab_migrate(Handle *h, bool carry) {
  const DevParams &p = h->p;
  SlabState &s = h->slab;
  DevArrays &a = h->a;
  const bool vel = h->second_kind == SECOND_VEL;
  const int peer[2] = {s.left, s.right};
  const size_t bytes = (size_t)a.em_cap * sizeof(EmRec);
  SLAB_NCCL(h, g_nccl.GroupStart());
  for (int d = 0; d < 2; d++) {
    SLAB_NCCL(h, g_nccl.Send(a.flags + FLAG_NEM0 + d, sizeof(int), ncclInt8, peer[d], s.comm, h->stream));
    SLAB_NCCL(h, g_nccl.Send(a.em_send[d], bytes, ncclInt8, peer[d], s.comm, h->stream));
  }
  // with 2 GPUs both messages come from the same peer: its first batch is what left through ITS left edge (towards me,
  // arriving on my right side), so the buffer of the right neighbour is received first
  for (int d = 1; d >= 0; d--) {
    SLAB_NCCL(h, g_nccl.Recv(a.flags + FLAG_NEMR0 + d, sizeof(int), ncclInt8, peer[d], s.comm, h->stream));
    SLAB_NCCL(h, g_nccl.Recv(a.em_recv[d], bytes, ncclInt8, peer[d], s.comm, h->stream));
  }
  SLAB_NCCL(h, g_nccl.GroupEnd());
  for (int d = 0; d < 2; d++) {
    k_ingest<<<8, 128, 0, h->stream>>>(p, a.em_recv[d], FLAG_NEMR0 + d, a.em_cap, vel ? 1 : 0, a.mv_pos, a.mv_second,
                                       a.mv_force, a.mv_id, a.mv_cell, a.mv_src, a.tile_dirty, a.dirty_list, a.inbox_cnt,
                                       a.inbox, a.flags, carry ? a.chg : nullptr);
    h->launches++;
  }
  return MAVI_OK;
}

// number of owned particles after a (re)build / migration -> p.n, p.n_active
static int slab_refresh_count(Handle *h) {
  ensure_rank_maps(h->ctx(), h->p, h->a);
  k_owned_count<<<1, 1, 0, h->stream>>>(h->p, h->a.tile_prefix, h->a.flags);
  h->launches++;
  int st = h->check_device_flags();
  if (st) return st;
  h->p.n = h->p.n_active = h->flags_host[FLAG_TAIL];
  h->p.n_count = h->slab.n_global;  // get_num_total_particles of the GLOBAL state (update_szabo!/update_rtp! loop 1:count)
  return MAVI_OK;
}

// Owned count + latched overflow word; called at the end of every mavi_step call (and every few steps inside it).
int slab_sync_counts(Handle *h) {
  const DevParams &p = h->p;
  int st = slab_refresh_count(h);
  if (st) return st;
  if (h->flags_host[FLAG_OVERFLOW]) {
    h->set_error("slab mode overflow (bits %d: 1 = tile capacity %d < %d, 2 = inbox capacity %d < %d at tile %d (column %d of %d), 4 = mover list %d, 8 = changed-cell list)",
                 h->flags_host[FLAG_OVERFLOW], p.cap, h->flags_host[FLAG_MAXCOUNT], p.inbox_cap, h->flags_host[FLAG_MAXINBOX],
                 h->flags_host[FLAG_MAXINBOX_TILE], h->flags_host[FLAG_MAXINBOX_TILE] / p.tpc, p.num_cols, p.mv_cap);
    return MAVI_ERR_CAPACITY;
  }
  return MAVI_OK;
}

// after the tile layout of the owned columns is final: counts + first halo exchange (positions and layout)
int slab_after_build(Handle *h) {
  int st = slab_refresh_count(h);
  if (st) return st;
  return slab_halo_exchange(h, h->a.pos[0], true);
}

// one step in slab mode (same operator order as Handle::step_once)
int slab_step_once(Handle *h, const double *noise_dev) {
  DevParams &p = h->p;
  DevArrays &a = h->a;
  LaunchCtx c = h->ctx();
  int st;
  if ((st = h->pending_out_of_grid())) return st;
  const bool vel = h->second_kind == SECOND_VEL;
  if (h->prof) cudaEventRecord(h->ev[0], h->stream);
  launch_step_begin(c, a);
  if (h->prof) cudaEventRecord(h->ev[1], h->stream);
  // force carry (see k_newton_b): F1 and the drift of this step were produced by the previous one
  const bool carry = vel && !(h->flags_cfg & MAVI_FLAG_NO_FORCE_CARRY);
  if (vel) {
    if (!carry || !h->carry_valid) launch_newton_a(c, p, a);      // owned: pos[0] -> pos[1], F1
    if (h->prof) cudaEventRecord(h->ev[2], h->stream);
    if ((st = slab_halo_exchange(h, a.pos[1], false))) return st; // drifted halo positions, same (stale) layout
    launch_newton_b(c, p, a, carry);
  } else {
    if (h->prof) cudaEventRecord(h->ev[2], h->stream);
    launch_self_propelled(c, p, a, noise_dev, (unsigned long long)h->num_steps);
  }
  std::swap(a.pos[0], a.pos[1]);
  if (h->prof) cudaEventRecord(h->ev[3], h->stream);
  // update_chunks! for the next step: emigrants -> neighbours, immigrants queued as movers, ONE incremental repair
  if ((st = slab_migrate(h, carry))) return st;
  launch_repair_tiles(c, p, a, vel);
  if (carry) launch_carry_redrift(c, p, a);
  h->time += p.dt;
  h->num_steps += 1;
  // (no host synchronisation here: the owned count and the overflow word are read by slab_sync_counts)
  if ((st = slab_halo_exchange(h, a.pos[0], true))) return st;
  if (carry) {  // needs the fresh halo (positions + layout)
    launch_carry_recompute(c, p, a);
    h->carry_valid = true;
  }
  if (h->prof) cudaEventRecord(h->ev[4], h->stream);
  return MAVI_OK;
}

}  // namespace mavi
