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
#include <mutex>
#include <utility>
#include <vector>

#include <cstdio>
#include <cstdlib>

#include "handle.cuh"

namespace MAVI_NS {

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
  const bool self = mp->world <= 1;  // MAVI_FLAG_SLAB_SELF: this rank is its own neighbour, no NCCL
  if (!self) {
    if (!mp->nccl_unique_id) {
      h->set_error("world > 1 needs MaviParams.nccl_unique_id (mavi_nccl_unique_id on rank 0, broadcast by the host)");
      return MAVI_ERR_BAD_PARAMS;
    }
    const char *e = load_nccl();
    if (e) {
      h->set_error("NCCL unavailable: %s", e);
      return MAVI_ERR_NCCL;
    }
  }
  s.world = self ? 1 : mp->world;
  s.n_global = (int)(mp->n_global > 0 ? mp->n_global : mp->n);
  s.rank = self ? 0 : mp->rank;
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
  if (!self) {
    ncclUniqueId id;
    memcpy(&id, mp->nccl_unique_id, sizeof id);
    {
      // The halo / migration messages are a few hundred KB: one or two channels carry them, and every channel is a CTA of
      // NCCL's send/recv kernel that needs (almost) an SM to itself.  NCCL's default here is 32 channels, i.e. a kernel that
      // wants 32 SMs next to a force pass that fills the device (2 GPUs: 0.671 ms/step with the default, 0.565 with <= 2
      // channels and four reserved SMs, kernels.cu grid_pipe()).  Only a default: the caller's environment wins; NCCL
      // reads the variable once per process, so a host that initialises NCCL itself first (bench.py under torchrun) sets it too.
      static std::once_flag nccl_env_once;
      std::call_once(nccl_env_once, [] { setenv("NCCL_MAX_P2P_NCHANNELS", "2", 0); });
    }
    SLAB_NCCL(h, g_nccl.CommInitRank(&s.comm, s.world, id, s.rank));
  }
  {  // the side stream's small kernels must not queue behind the 16k blocks of the interior pass
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    SLAB_CUDA(h, cudaStreamCreateWithPriority(&s.side, cudaStreamNonBlocking, hi));
  }
  SLAB_CUDA(h, cudaEventCreateWithFlags(&s.ev_ready, cudaEventDisableTiming));
  SLAB_CUDA(h, cudaEventCreateWithFlags(&s.ev_halo, cudaEventDisableTiming));
  SLAB_CUDA(h, cudaEventCreateWithFlags(&s.ev_begin, cudaEventDisableTiming));
  SLAB_CUDA(h, cudaEventCreateWithFlags(&s.ev_side, cudaEventDisableTiming));
  return MAVI_OK;
}

// max over all ranks of one int (tile capacity agreement: every rank must use the SAME slots-per-column, because a
// halo / emigrant column is shipped as one raw slab of tpc*cap slots)
int slab_allreduce_max(Handle *h, int *value) {
  if (!h->slab.comm) return MAVI_OK;  // one-rank slab mode
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
  if (h->slab.side) cudaStreamDestroy(h->slab.side);
  if (h->slab.ev_ready) cudaEventDestroy(h->slab.ev_ready);
  if (h->slab.ev_halo) cudaEventDestroy(h->slab.ev_halo);
  if (h->slab.ev_begin) cudaEventDestroy(h->slab.ev_begin);
  if (h->slab.ev_side) cudaEventDestroy(h->slab.ev_side);
  h->slab.ev_begin = h->slab.ev_side = nullptr;
  h->slab.side_busy = false;
  h->slab.side = nullptr;
  h->slab.ev_ready = h->slab.ev_halo = nullptr;
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
                         int has_vel, real2 *__restrict__ mv_pos, real2 *__restrict__ mv_second,
                         real2 *__restrict__ mv_force, unsigned int *__restrict__ mv_id, int *__restrict__ mv_cell,
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
    raise_mark(&flags[FLAG_INBOX_STEP], i + 1);
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
int slab_halo_exchange(Handle *h, real2 *pos_buf, bool with_layout, cudaStream_t stream = nullptr) {
  if (!stream) stream = h->stream;
  const DevParams &p = h->p;
  SlabState &s = h->slab;
  DevArrays &a = h->a;
  const size_t cs = col_slots(p), m = (size_t)s.m;
  const size_t rows = (size_t)p.tpc * TR1;
  // my boundary columns: local 1 (-> left neighbour's right halo) and local m (-> right neighbour's left halo)
  if (!s.comm) {  // one-rank slab mode: I am my own left and right neighbour
    SLAB_CUDA(h, cudaMemcpyAsync(pos_buf + (m + 1) * cs, pos_buf + 1 * cs, cs * sizeof(real2), cudaMemcpyDeviceToDevice, stream));
    SLAB_CUDA(h, cudaMemcpyAsync(pos_buf + 0 * cs, pos_buf + m * cs, cs * sizeof(real2), cudaMemcpyDeviceToDevice, stream));
    if (with_layout) {
      SLAB_CUDA(h, cudaMemcpyAsync(a.tstart + (m + 1) * rows, a.tstart + 1 * rows, rows * sizeof(int), cudaMemcpyDeviceToDevice, stream));
      SLAB_CUDA(h, cudaMemcpyAsync(a.tstart + 0 * rows, a.tstart + m * rows, rows * sizeof(int), cudaMemcpyDeviceToDevice, stream));
    }
  } else {
  SLAB_NCCL(h, g_nccl.GroupStart());
  SLAB_NCCL(h, g_nccl.Send(pos_buf + 1 * cs, cs * sizeof(real2), ncclInt8, s.left, s.comm, stream));
  SLAB_NCCL(h, g_nccl.Send(pos_buf + m * cs, cs * sizeof(real2), ncclInt8, s.right, s.comm, stream));
  // receive order matters when left == right (2 GPUs): the peer's FIRST send is its column 1 = my RIGHT halo
  SLAB_NCCL(h, g_nccl.Recv(pos_buf + (m + 1) * cs, cs * sizeof(real2), ncclInt8, s.right, s.comm, stream));
  SLAB_NCCL(h, g_nccl.Recv(pos_buf + 0 * cs, cs * sizeof(real2), ncclInt8, s.left, s.comm, stream));
  if (with_layout) {
    SLAB_NCCL(h, g_nccl.Send(a.tstart + 1 * rows, rows * sizeof(int), ncclInt8, s.left, s.comm, stream));
    SLAB_NCCL(h, g_nccl.Send(a.tstart + m * rows, rows * sizeof(int), ncclInt8, s.right, s.comm, stream));
    SLAB_NCCL(h, g_nccl.Recv(a.tstart + (m + 1) * rows, rows * sizeof(int), ncclInt8, s.right, s.comm, stream));
    SLAB_NCCL(h, g_nccl.Recv(a.tstart + 0 * rows, rows * sizeof(int), ncclInt8, s.left, s.comm, stream));
  }
  SLAB_NCCL(h, g_nccl.GroupEnd());
  }
  if (with_layout) {
    // right halo came from the right neighbour's column 1; left halo from the left neighbour's column m_left
    const int nb = ((int)rows + 255) / 256;
    k_rebase_tstart<<<nb, 256, 0, stream>>>(a.tstart + (m + 1) * rows, (int)rows, (int)((m + 1) * cs) - (int)(1 * cs));
    k_rebase_tstart<<<nb, 256, 0, stream>>>(a.tstart + 0 * rows, (int)rows, 0 - (int)((size_t)s.m_left * cs));
    h->launches += 2;
  }
  return MAVI_OK;
}

// Migration: the integrate kernel wrote the record of every particle that crossed into a halo column straight into
// the emigrant buffer of that side (note_if_moved).  Ship both buffers (fixed capacity, the count travels separately)
// and queue what arrived as movers; the ONE tile repair of the step then handles local movers and immigrants alike.
static int slab_migrate(Handle *h, bool carry, cudaStream_t stream) {
  const DevParams &p = h->p;
  SlabState &s = h->slab;
  DevArrays &a = h->a;
  const bool vel = h->second_kind == SECOND_VEL;
  const int peer[2] = {s.left, s.right};
  const size_t bytes = (size_t)a.em_cap * sizeof(EmRec);
  if (!s.comm) {  // one-rank slab mode: what leaves through my left edge arrives from my right, and vice versa
    for (int d = 0; d < 2; d++) {
      SLAB_CUDA(h, cudaMemcpyAsync(a.flags + FLAG_NEMR0 + (1 - d), a.flags + FLAG_NEM0 + d, sizeof(int), cudaMemcpyDeviceToDevice, stream));
      SLAB_CUDA(h, cudaMemcpyAsync(a.em_recv[1 - d], a.em_send[d], bytes, cudaMemcpyDeviceToDevice, stream));
    }
  } else {
  SLAB_NCCL(h, g_nccl.GroupStart());
  for (int d = 0; d < 2; d++) {
    SLAB_NCCL(h, g_nccl.Send(a.flags + FLAG_NEM0 + d, sizeof(int), ncclInt8, peer[d], s.comm, stream));
    SLAB_NCCL(h, g_nccl.Send(a.em_send[d], bytes, ncclInt8, peer[d], s.comm, stream));
  }
  // with 2 GPUs both messages come from the same peer: its first batch is what left through ITS left edge (towards me,
  // arriving on my right side), so the buffer of the right neighbour is received first
  for (int d = 1; d >= 0; d--) {
    SLAB_NCCL(h, g_nccl.Recv(a.flags + FLAG_NEMR0 + d, sizeof(int), ncclInt8, peer[d], s.comm, stream));
    SLAB_NCCL(h, g_nccl.Recv(a.em_recv[d], bytes, ncclInt8, peer[d], s.comm, stream));
  }
  SLAB_NCCL(h, g_nccl.GroupEnd());
  }
  for (int d = 0; d < 2; d++) {
    k_ingest<<<8, 128, 0, stream>>>(p, a.em_recv[d], FLAG_NEMR0 + d, a.em_cap, vel ? 1 : 0, a.mv_pos, a.mv_second,
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
int slab_join(Handle *h);
int slab_sync_counts(Handle *h) {
  const DevParams &p = h->p;
  int st = slab_join(h);
  if (st) return st;
  st = slab_refresh_count(h);
  if (st) return st;
  // Proactive tile growth.  A tile overflow in the middle of an un-synchronised run of slab steps cannot be rolled back (the
  // other ranks have moved on), so it must not happen: at every synchronisation point (all ranks are at the same step here)
  // the fullest tile seen so far is compared with the capacity and, if ANY rank is above 85 %, every rank rebuilds its
  // layout with larger tiles — the slab-mode counterpart of the single-GPU overflow -> rebuild -> resume path.
  if (!h->flags_host[FLAG_OVERFLOW]) {
    int need = 0;
    // (a tile of the benchmark lattice holds up to 72 particles of the initial capacity 96: 75 % full is normal)
    if (h->flags_host[FLAG_MAXCOUNT] * 100 > p.cap * 85) need = ((int)(h->flags_host[FLAG_MAXCOUNT] * 1.3) + 15) / 16 * 16;
    if ((st = slab_allreduce_max(h, &need))) return st;
    if (need > p.cap) {
      h->n_rebuilds++;
      launch_compact_to_staging(h->ctx(), h->p, h->a, h->second_kind == SECOND_VEL);  // with the current layout
      const int n_owned = h->p.n;
      if ((st = h->alloc_state(n_owned, need))) return st;          // every rank: the same, all-reduced capacity
      if ((st = h->rebuild_from_staging(n_owned))) return st;       // collective (capacity agreement, first halo exchange)
      return MAVI_OK;
    }
  }
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
// MAVI_SLAB_TRACE=<step>: per-phase device timeline of that step on stderr (debugging aid; synchronises once)
struct SlabTrace {
  static constexpr int N = 12;
  cudaEvent_t ev[N];
  bool on = false;
  bool rec[N] = {};
  void begin() { for (auto &e : ev) cudaEventCreate(&e); on = true; }
  void mark(int i, cudaStream_t s) { if (on) { cudaEventRecord(ev[i], s); rec[i] = true; } }
  void end(int rank, cudaStream_t main, cudaStream_t side, const int *flags_dev, const DevParams &p) {
    if (!on) return;
    cudaStreamSynchronize(main);
    cudaStreamSynchronize(side);
    int f[FLAG_COUNT];
    cudaMemcpy(f, flags_dev, sizeof f, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[slab trace rank %d] flags: err=%d dirty=%d bigmove=%d nfix=%d nmv=%d ran=%d overflow=%d nchg=%d bigmove_next=%d nem=%d,%d nemr=%d,%d | n=%d cap=%d blk_cols=%d bpr=%d cols=%d fast_interior=%d\n",
            rank, f[FLAG_ERR], f[FLAG_CHANGED], f[FLAG_BIGMOVE], f[FLAG_NFIX], f[FLAG_NMV], f[FLAG_RAN], f[FLAG_OVERFLOW], f[FLAG_NCHG],
            f[FLAG_BIGMOVE_NEXT], f[FLAG_NEM0], f[FLAG_NEM1], f[FLAG_NEMR0], f[FLAG_NEMR1], p.n, p.cap, p.blk_cols, p.blk_per_row,
            p.num_cols, p.fast_interior);
    const char *names[N] = {"start", "step_begin", "K_int done", "fixes done (after side wait)", "migrate+ingest", "repair+redrift",
                            "side: C (halo X+layout) done", "interior recompute done", "side: K_bnd start", "side: bnd recompute + A(next) done", "side: K_bnd done", ""};
    for (int i = 1; i < 11; i++) {
      if (!rec[i]) continue;
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[0], ev[i]);
      fprintf(stderr, "[slab trace rank %d] %-32s %8.1f us\n", rank, names[i], ms * 1e3f);
    }
    for (auto &e : ev) cudaEventDestroy(e);
    on = false;
  }
};

// Joins the side stream into the main one (before anything that is not part of the step loop touches the state).
int slab_join(Handle *h) {
  if (!h->slab.side_busy) return MAVI_OK;
  SLAB_CUDA(h, cudaEventRecord(h->slab.ev_side, h->slab.side));
  SLAB_CUDA(h, cudaStreamWaitEvent(h->stream, h->slab.ev_side, 0));
  return MAVI_OK;
}

// One step in slab mode.
//
// Newton steps with the force carry are pipelined over two streams.  Per step n, in steady state:
//   side : ... [C(n-1): halo positions + layout] [boundary recompute(n-1)] [A(n): drifted halo]  K_bnd(n)  B(n): emigrant
//          records + ingest
//   main : step_begin(n)  K_int(n)  | wait side |  wall fix-ups, tile repair, re-drift, list-driven recompute of the interior
// K_int = the blocks that read no halo column, K_bnd = the first / last block of every tile row.  The interior work of
// step n+1 never touches the two owned columns next to a halo, so the exchanges C and A and the recomputation of those
// columns (always done in full: the neighbour's boundary column re-bins behind this rank's back) overlap with K_int.
// Communicator operations all live on the side stream (A(n) -> B(n) -> C(n) -> A(n+1)), identically ordered on all ranks.
int slab_step_once(Handle *h, const real *noise_dev) {
  DevParams &p = h->p;
  DevArrays &a = h->a;
  SlabState &s = h->slab;
  LaunchCtx c = h->ctx();
  LaunchCtx cside = c;
  cside.stream = s.side;
  int st;
  static const long long trace_step = getenv("MAVI_SLAB_TRACE") ? atoll(getenv("MAVI_SLAB_TRACE")) : -1;
  SlabTrace tr;
  if (h->num_steps == trace_step) { cudaStreamSynchronize(h->stream); cudaStreamSynchronize(s.side); tr.begin(); }
  tr.mark(0, h->stream);
  if ((st = h->pending_out_of_grid())) return st;
  const bool vel = h->second_kind == SECOND_VEL;
  // force carry (see k_newton_b): F1 and the drift of this step were produced by the previous one
  const bool carry = vel && !(h->flags_cfg & MAVI_FLAG_NO_FORCE_CARRY);
  static const bool nopipe = getenv("MAVI_SLAB_NOPIPE") != nullptr;  // debugging aid: everything on the main stream
  const bool piped = carry && !nopipe && s.m >= 2 * MAVI_EDGE_COLS + 2;
  if (h->prof) cudaEventRecord(h->ev[0], h->stream);
  launch_step_begin(c, a);
  tr.mark(1, h->stream);
  if (h->prof) cudaEventRecord(h->ev[1], h->stream);
  if (piped) {
    if (!h->carry_valid) {
      // prime the pipeline: full first pass, then the drifted halo on the side stream
      if ((st = slab_join(h))) return st;
      launch_newton_a(c, p, a);
      SLAB_CUDA(h, cudaEventRecord(s.ev_ready, h->stream));
      SLAB_CUDA(h, cudaStreamWaitEvent(s.side, s.ev_ready, 0));
      if ((st = slab_halo_exchange(h, a.pos[1], false, s.side))) return st;
      s.side_busy = true;
    }
    if (h->prof) cudaEventRecord(h->ev[2], h->stream);
    SLAB_CUDA(h, cudaEventRecord(s.ev_begin, h->stream));  // FLAG_RAN and the per-step counters of this step are set
    SLAB_CUDA(h, cudaStreamWaitEvent(s.side, s.ev_begin, 0));
    tr.mark(8, s.side);
    launch_newton_b(cside, p, a, true, 2);  // blocks next to a halo column, after A on the side stream
    tr.mark(10, s.side);
    // B: every emigrant comes out of a boundary block (a particle of an interior block would have to cross >= 3 cell
    // columns in one step; note_moved_slow reports that loudly), so the migration exchange and the ingestion of the
    // immigrants run on the side stream right away, overlapped with the interior blocks.  k_ingest only appends to the
    // mover / dirty / changed-cell lists with the same atomics the interior blocks use.
    if ((st = slab_migrate(h, carry, s.side))) return st;
    tr.mark(4, s.side);
    SLAB_CUDA(h, cudaEventRecord(s.ev_halo, s.side));
    launch_newton_b(c, p, a, true, 1);
    tr.mark(2, h->stream);
    SLAB_CUDA(h, cudaStreamWaitEvent(h->stream, s.ev_halo, 0));
    launch_apply_pos_fixes(c, a);
    tr.mark(3, h->stream);
  } else if (vel) {
    if (!carry || !h->carry_valid) launch_newton_a(c, p, a);      // owned: pos[0] -> pos[1], F1
    if (h->prof) cudaEventRecord(h->ev[2], h->stream);
    if ((st = slab_halo_exchange(h, a.pos[1], false))) return st; // drifted halo positions, same (stale) layout
    tr.mark(8, h->stream);
    launch_newton_b(c, p, a, carry);
    tr.mark(3, h->stream);
  } else {
    if (h->prof) cudaEventRecord(h->ev[2], h->stream);
    launch_self_propelled(c, p, a, noise_dev, (unsigned long long)h->num_steps);
  }
  std::swap(a.pos[0], a.pos[1]);
  if (h->prof) cudaEventRecord(h->ev[3], h->stream);
  // update_chunks! for the next step: emigrants -> neighbours, immigrants queued as movers, ONE incremental repair
  if (!piped) {
    if ((st = slab_migrate(h, carry, h->stream))) return st;
    tr.mark(4, h->stream);
  }
  launch_repair_tiles(c, p, a, vel);
  if (carry) launch_carry_redrift(c, p, a);
  tr.mark(5, h->stream);
  h->time += h->dt_host;
  h->num_steps += 1;
  // (no host synchronisation here: the owned count and the overflow word are read by slab_sync_counts)
  if (piped) {
    SLAB_CUDA(h, cudaEventRecord(s.ev_ready, h->stream));
    launch_carry_recompute_list(c, p, a, 3);  // interior: leaves the halo and the two owned columns next to it alone
    tr.mark(7, h->stream);
    SLAB_CUDA(h, cudaStreamWaitEvent(s.side, s.ev_ready, 0));
    if ((st = slab_halo_exchange(h, a.pos[0], true, s.side))) return st;  // C: fresh halo positions + layout
    tr.mark(6, s.side);
    launch_carry_recompute_columns(cside, p, a, 2, false);
    if ((st = slab_halo_exchange(h, a.pos[1], false, s.side))) return st;  // A of the next step: drifted halo
    tr.mark(9, s.side);
    s.side_busy = true;
    h->carry_valid = true;
  } else {
    if ((st = slab_halo_exchange(h, a.pos[0], true))) return st;
    tr.mark(6, h->stream);
    if (carry) {  // needs the fresh halo (positions + layout)
      launch_carry_recompute(c, p, a);
      h->carry_valid = true;
    }
    tr.mark(7, h->stream);
  }
  if (h->prof) cudaEventRecord(h->ev[4], h->stream);
  tr.end(s.rank, h->stream, s.side, a.flags, p);
  return MAVI_OK;
}

}  // namespace MAVI_NS
