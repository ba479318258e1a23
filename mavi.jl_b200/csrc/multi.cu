// multi.cu — single-process multi-GPU handle (MaviParams.n_gpus > 1).
//
// The reference drives ONE System from ONE process (src/run_system.jl:7-23, src/systems.jl:73-114); its Threaded mode
// splits the pair loop over cell columns across threads (src/integration.jl:159-194).  Here the same column split goes
// over the GPUs of the node INSIDE the handle: mavi_create(n_gpus = G) makes one x-slab sub-handle per device (slab.cu,
// the machinery bench.py drives with one process per GPU) and one host worker thread per device; NCCL communicators come
// from ncclCommInitRank called concurrently by the workers.  The caller keeps the plain single-GPU calls:
//   mavi_upload_state    partitions by cell column with the device's exact cell rule (hostcell.h) and uploads every slab
//   mavi_step            runs the slab step loop on every device concurrently (halo + migration over NVLink)
//   mavi_download_state / _forces / _cells   gather and un-permute to the caller's original ids
// Nothing here depends on the arithmetic type: the per-dtype entry points are reached through a function table.
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <numeric>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mavi.h"
#include "hostcell.h"
#include "multi.h"

namespace mavi_multi {

// ---- a fixed pool of one worker thread per device ---------------------------------------------------------------------
class Pool {
 public:
  explicit Pool(int n) : n_(n), status_(n, 0) {
    for (int g = 0; g < n; g++) threads_.emplace_back([this, g] { loop(g); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto &t : threads_) t.join();
  }
  // fn(g) on every worker; returns the first non-zero status (in rank order)
  int run(const std::function<int(int)> &fn) {
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      pending_ = n_;
      gen_++;
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
    for (int g = 0; g < n_; g++)
      if (status_[g]) return status_[g];
    return 0;
  }
  int status(int g) const { return status_[g]; }

 private:
  void loop(int g) {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<int(int)> *fn;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        fn = fn_;
      }
      const int st = (*fn)(g);
      {
        std::lock_guard<std::mutex> lk(m_);
        status_[g] = st;
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  int n_;
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<int(int)> *fn_ = nullptr;
  unsigned long long gen_ = 0;
  int pending_ = 0;
  bool stop_ = false;
  std::vector<int> status_;
};

// pinned host staging of one rank (grown on demand)
struct HostBuf {
  void *ptr = nullptr;
  size_t cap = 0;
  void *get(size_t bytes) {
    if (bytes <= cap) return ptr;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    if (cudaMallocHost(&ptr, want) != cudaSuccess) {
      cudaGetLastError();
      ptr = nullptr;
      return nullptr;
    }
    cap = want;
    return ptr;
  }
  ~HostBuf() {
    if (ptr) cudaFreeHost(ptr);
  }
};

struct Rank {
  void *impl = nullptr;
  int device = 0;
  int col_lo = 0, m = 0;
  int64_t n_local = 0;
  int64_t cap = 0;           // dense-array capacity of the sub-handle (api.cu allocate: 1.25 * hint + 4096)
  std::vector<int64_t> idx;  // upload: original ids routed to this rank
  HostBuf ids, pos, second, force, cells;
};

struct Multi {
  const ApiTable *api = nullptr;
  MaviParams params;  // the caller's block (n = global particle count); pointers inside are never dereferenced later
  int G = 0;
  size_t elem = 8;          // sizeof(T)
  int second_width = 2;     // vel: 2 reals per particle, pol_angle: 1
  std::vector<Rank> ranks;
  Pool *pool = nullptr;
  bool created = false;
  int64_t num_steps0 = 0;
  double time0 = 0.0;
  bool prof = false;
  char err[512] = {0};
  mavi_host::Grid grid;

  void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, sizeof err, fmt, ap);
    va_end(ap);
  }
  void take_rank_error(int g) {
    char buf[400] = {0};
    if (ranks[g].impl) api->last_error(ranks[g].impl, buf, (int32_t)sizeof buf);
    set_error("GPU %d (slab %d of %d): %s", ranks[g].device, g, G, buf);
  }
  int first_error(int st) {
    if (!st) return 0;
    for (int g = 0; g < G; g++)
      if (pool->status(g)) {
        take_rank_error(g);
        break;
      }
    return st;
  }
  void destroy_ranks() {
    if (!created) return;
    pool->run([&](int g) {
      if (ranks[g].impl) api->destroy(ranks[g].impl);
      ranks[g].impl = nullptr;
      return 0;
    });
    created = false;
  }
};

static void slab_columns(int C, int world, int rank, int *lo, int *m) {  // == slab.cu
  const int base = C / world, rem = C % world;
  *m = base + (rank < rem ? 1 : 0);
  *lo = rank * base + (rank < rem ? rank : rem);
}
static int owner_of_column(int col, int C, int world) {
  const int base = C / world, rem = C % world;
  const int split = rem * (base + 1);
  return col < split ? col / (base + 1) : rem + (col - split) / (base > 0 ? base : 1);
}

static double coord(const void *pos, size_t elem, int64_t i, int d) {
  return elem == 4 ? (double)((const float *)pos)[2 * i + d] : ((const double *)pos)[2 * i + d];
}

// ---- lifetime -----------------------------------------------------------------------------------------------------------
int create(const ApiTable *api, const MaviParams *p, void **out) {
  Multi *m = new Multi();
  *out = m;
  m->api = api;
  m->params = *p;
  m->G = p->n_gpus;
  m->elem = p->dtype == MAVI_F32 ? 4 : 8;
  m->second_width = (p->dynamics == MAVI_DYN_LJ || p->dynamics == MAVI_DYN_HARMTRUNC) ? 2 : 1;
  if (p->struct_size != sizeof(MaviParams)) {
    m->set_error("MaviParams.struct_size %u != %zu (ABI mismatch)", p->struct_size, sizeof(MaviParams));
    return MAVI_ERR_BAD_PARAMS;
  }
  if (p->world > 1) {
    m->set_error("n_gpus > 1 (one process, all GPUs inside the handle) and world > 1 (one process per GPU) are exclusive");
    return MAVI_ERR_BAD_PARAMS;
  }
  if (p->n_spaces != 1 || p->spaces[0].wall != MAVI_WALL_PERIODIC || p->spaces[0].geom != MAVI_GEOM_RECT || p->num_cols <= 0 ||
      p->dynamics == MAVI_DYN_RINGS) {
    m->set_error("n_gpus > 1 needs a single periodic rectangle with chunks and a particle (not Rings) dynamics: all-pairs "
                 "mode, composite spaces and Mavi.Rings run on one GPU (SURVEY.md 8e: replicas only)");
    return MAVI_ERR_UNSUPPORTED;
  }
  if (p->num_cols / p->n_gpus < 2) {
    m->set_error("every slab needs at least 2 cell columns (%d columns over %d GPUs)", p->num_cols, p->n_gpus);
    return MAVI_ERR_BAD_PARAMS;
  }
  if (p->stream) {
    m->set_error("n_gpus > 1: MaviParams.stream must be NULL (a stream belongs to one device)");
    return MAVI_ERR_BAD_PARAMS;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < p->device + p->n_gpus) {
    cudaGetLastError();
    m->set_error("n_gpus = %d starting at device %d, but %d CUDA devices are visible — there is no CPU fallback", p->n_gpus,
                 p->device, ndev);
    return MAVI_ERR_CUDA;
  }
  m->grid.bl[0] = p->grid_bl[0];
  m->grid.bl[1] = p->grid_bl[1];
  m->grid.h = p->grid_h;
  m->grid.cl = p->grid_len / (double)p->num_cols;
  m->grid.ch = p->grid_h / (double)p->num_rows;
  m->grid.cols = p->num_cols;
  m->grid.rows = p->num_rows;
  m->ranks.resize(m->G);
  for (int g = 0; g < m->G; g++) {
    m->ranks[g].device = p->device + g;
    slab_columns(p->num_cols, m->G, g, &m->ranks[g].col_lo, &m->ranks[g].m);
  }
  m->pool = new Pool(m->G);
  return MAVI_OK;
}

int destroy(void *mm) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m) return MAVI_OK;
  if (m->pool) {
    m->destroy_ranks();
    delete m->pool;
  }
  delete m;
  return MAVI_OK;
}

int last_error(void *mm, char *buf, int32_t n) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m || !buf || n <= 0) return MAVI_ERR_BAD_PARAMS;
  snprintf(buf, (size_t)n, "%s", m->err);
  return MAVI_OK;
}

// one slab sub-handle per device, sized for the partition at hand
static int create_ranks(Multi *m) {
  m->destroy_ranks();
  unsigned char uid[128];
  int st = m->api->nccl_unique_id(uid);
  if (st) {
    m->set_error("NCCL unavailable (libnccl.so.2 could not be loaded)");
    return st;
  }
  st = m->pool->run([&](int g) {
    MaviParams q = m->params;
    q.n_gpus = 0;
    q.rank = g;
    q.world = m->G;
    q.device = m->ranks[g].device;
    q.nccl_unique_id = uid;
    q.n_global = m->params.n;
    // capacity hint: the dense arrays get 25 % + 4096 head room over this (api.cu allocate)
    int64_t hint = std::max<int64_t>(m->ranks[g].n_local, m->params.n / m->G);
    q.n = hint;
    m->ranks[g].cap = (int64_t)((int)(hint * 1.25) + 4096);
    return (int)m->api->create(&q, &m->ranks[g].impl);
  });
  m->created = true;  // sub-handles exist (possibly in an error state): destroy_ranks frees them
  if (st) return m->first_error(st);
  st = m->pool->run([&](int g) {
    int s = m->api->set_time(m->ranks[g].impl, m->num_steps0, m->time0);
    if (!s && m->prof) s = m->api->set_profiling(m->ranks[g].impl, 1);
    return s;
  });
  return m->first_error(st);
}

// ---- state movement -------------------------------------------------------------------------------------------------------
int upload_state(void *mm, const void *pos, const void *second, const uint8_t *mask, int64_t n) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m || !pos || n != m->params.n) return MAVI_ERR_BAD_PARAMS;
  if (mask)
    for (int64_t i = 0; i < n; i++)
      if (!mask[i]) {
        m->set_error("n_gpus > 1: inactive particle slots (ParticleIds masks) are single-GPU only");
        return MAVI_ERR_UNSUPPORTED;
      }
  const int G = m->G, C = m->params.num_cols;
  // owner of every particle: cell column by the device's exact rule (update_particle_chunk!, src/chunks.jl:120-147)
  std::vector<int32_t> owner((size_t)n);
  std::atomic<int> bad{0};
  m->pool->run([&](int g) {
    const int64_t lo = n * g / G, hi = n * (g + 1) / G;
    for (int64_t i = lo; i < hi; i++) {
      int col, row;
      if (!mavi_host::cell_of_point(m->grid, coord(pos, m->elem, i, 0), coord(pos, m->elem, i, 1), &col, &row)) {
        bad.store(1);
        owner[(size_t)i] = -1;
      } else {
        owner[(size_t)i] = owner_of_column(col, C, G);
      }
    }
    return 0;
  });
  if (bad.load()) {
    m->set_error("a particle lies outside the chunk grid (BoundsError in the reference, src/chunks.jl:144-146)");
    return MAVI_ERR_OUT_OF_GRID;
  }
  bool fits = m->created;
  m->pool->run([&](int g) {
    Rank &r = m->ranks[g];
    r.idx.clear();
    for (int64_t i = 0; i < n; i++)
      if (owner[(size_t)i] == g) r.idx.push_back(i);
    r.n_local = (int64_t)r.idx.size();
    return 0;
  });
  std::vector<int32_t>().swap(owner);
  // a slab that outgrew the dense arrays of its sub-handle: re-create ALL sub-handles before anyone enters a collective
  for (auto &r : m->ranks)
    if (r.n_local > r.cap) fits = false;
  if (!fits) {
    int st = create_ranks(m);
    if (st) return st;
  }
  const size_t e = m->elem, sw = (size_t)m->second_width;
  {
    int st = m->pool->run([&](int g) {
      Rank &r = m->ranks[g];
      const size_t nl = (size_t)r.n_local;
      int64_t *ids = (int64_t *)r.ids.get((nl + 1) * sizeof(int64_t));
      char *p = (char *)r.pos.get((nl + 1) * 2 * e);
      char *s = second ? (char *)r.second.get((nl + 1) * sw * e) : nullptr;
      if (!ids || !p || (second && !s)) return (int)MAVI_ERR_CUDA;
      for (size_t k = 0; k < nl; k++) {
        const int64_t i = r.idx[k];
        ids[k] = i;
        memcpy(p + k * 2 * e, (const char *)pos + (size_t)i * 2 * e, 2 * e);
        if (s) memcpy(s + k * sw * e, (const char *)second + (size_t)i * sw * e, sw * e);
      }
      return (int)m->api->upload_local(r.impl, ids, p, s, (int64_t)nl);
    });
    if (st) return m->first_error(st);
  }
  for (auto &r : m->ranks) std::vector<int64_t>().swap(r.idx);
  return MAVI_OK;
}

// every rank downloads what it owns into pinned staging and scatters it to the caller's arrays by original id
static int gather(Multi *m, void *pos, void *second, void *forces, int32_t *cells) {
  if (!m->created) {
    m->set_error("no state uploaded yet");
    return MAVI_ERR_BAD_PARAMS;
  }
  const size_t e = m->elem, sw = (size_t)m->second_width;
  int st = m->pool->run([&](int g) {
    Rank &r = m->ranks[g];
    int64_t nl = 0;
    int s = m->api->local_count(r.impl, &nl);
    if (s) return s;
    r.n_local = nl;
    const size_t n1 = (size_t)nl + 1;
    int64_t *ids = (int64_t *)r.ids.get(n1 * sizeof(int64_t));
    char *p = pos ? (char *)r.pos.get(n1 * 2 * e) : nullptr;
    char *sc = second ? (char *)r.second.get(n1 * sw * e) : nullptr;
    char *f = forces ? (char *)r.force.get(n1 * 2 * e) : nullptr;
    int32_t *cl = cells ? (int32_t *)r.cells.get(n1 * sizeof(int32_t)) : nullptr;
    if (!ids || (pos && !p) || (second && !sc) || (forces && !f) || (cells && !cl)) return (int)MAVI_ERR_CUDA;
    s = m->api->download_local(r.impl, ids, p, sc, f);
    if (s) return s;
    if (cl && (s = m->api->download_local_cells(r.impl, cl))) return s;
    for (size_t k = 0; k < (size_t)nl; k++) {
      const size_t i = (size_t)ids[k];
      if (p) memcpy((char *)pos + i * 2 * e, p + k * 2 * e, 2 * e);
      if (sc) memcpy((char *)second + i * sw * e, sc + k * sw * e, sw * e);
      if (f) memcpy((char *)forces + i * 2 * e, f + k * 2 * e, 2 * e);
      if (cl) cells[i] = cl[k];
    }
    return 0;
  });
  if (st) return m->first_error(st);
  int64_t total = 0;
  for (auto &r : m->ranks) total += r.n_local;
  if (total != m->params.n) {
    m->set_error("the slabs own %lld particles, the state has %lld", (long long)total, (long long)m->params.n);
    return MAVI_ERR_CAPACITY;
  }
  return MAVI_OK;
}

int download_state(void *mm, void *pos, void *second) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m) return MAVI_ERR_BAD_PARAMS;
  if (!pos && !second) return MAVI_OK;
  return gather(m, pos, second, nullptr, nullptr);
}

int download_forces(void *mm, void *forces) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m || !forces) return MAVI_ERR_BAD_PARAMS;
  return gather(m, nullptr, nullptr, forces, nullptr);
}

int local_count(void *mm, int64_t *n) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m || !n) return MAVI_ERR_BAD_PARAMS;
  *n = m->params.n;
  return MAVI_OK;
}

int download_local(void *mm, int64_t *ids, void *pos, void *second, void *forces) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m) return MAVI_ERR_BAD_PARAMS;
  if (ids)
    for (int64_t i = 0; i < m->params.n; i++) ids[i] = i;
  if (!pos && !second && !forces) return MAVI_OK;
  return gather(m, pos, second, forces, nullptr);
}

int upload_local(void *mm, const int64_t *, const void *, const void *, int64_t) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m) return MAVI_ERR_BAD_PARAMS;
  m->set_error("n_gpus > 1 partitions inside mavi_upload_state; mavi_upload_local is the one-process-per-GPU entry point");
  return MAVI_ERR_BAD_PARAMS;
}

// ---- hot path ---------------------------------------------------------------------------------------------------------------
#define MULTI_ALL(m, call)                                                 \
  do {                                                                     \
    if (!(m)) return MAVI_ERR_BAD_PARAMS;                                  \
    if (!(m)->created) {                                                   \
      (m)->set_error("no state uploaded yet");                             \
      return MAVI_ERR_BAD_PARAMS;                                          \
    }                                                                      \
    int st_ = (m)->pool->run([&](int g) { return (int)(m)->api->call; });  \
    return (m)->first_error(st_);                                          \
  } while (0)

int step(void *mm, int64_t nsteps, const void *host_noise) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  // host noise rows are indexed by the original (global) id: every slab gets the same rows
  MULTI_ALL(m, step(m->ranks[g].impl, nsteps, host_noise));
}
int calc_forces(void *mm) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  MULTI_ALL(m, calc_forces(m->ranks[g].impl));
}
int bin(void *mm) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  MULTI_ALL(m, bin(m->ranks[g].impl));
}
int sync(void *mm) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  MULTI_ALL(m, sync(m->ranks[g].impl));
}

int set_profiling(void *mm, int32_t on) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m) return MAVI_ERR_BAD_PARAMS;
  m->prof = on != 0;
  if (!m->created) return MAVI_OK;
  MULTI_ALL(m, set_profiling(m->ranks[g].impl, on));
}

int download_cells(void *mm, int32_t *cell_of_particle, int32_t *counts) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m) return MAVI_ERR_BAD_PARAMS;
  const int64_t n = m->params.n;
  std::vector<int32_t> tmp;
  int32_t *cells = cell_of_particle;
  if (!cells) {
    tmp.resize((size_t)n);
    cells = tmp.data();
  }
  int st = gather(m, nullptr, nullptr, nullptr, cells);
  if (st) return st;
  if (counts) {
    const int64_t nc = (int64_t)m->params.num_cols * m->params.num_rows;
    std::fill(counts, counts + nc, 0);
    for (int64_t i = 0; i < n; i++)
      if (cells[i] >= 0) counts[cells[i]]++;
  }
  return MAVI_OK;
}

int download_cell_lists(void *mm, int32_t *start, int32_t *ids) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m) return MAVI_ERR_BAD_PARAMS;
  const int64_t n = m->params.n, nc = (int64_t)m->params.num_cols * m->params.num_rows;
  std::vector<int32_t> cells((size_t)n), cnt((size_t)nc + 1, 0);
  int st = gather(m, nullptr, nullptr, nullptr, cells.data());
  if (st) return st;
  for (int64_t i = 0; i < n; i++) cnt[(size_t)cells[i] + 1]++;
  for (int64_t c = 0; c < nc; c++) cnt[(size_t)c + 1] += cnt[(size_t)c];
  if (start) std::copy(cnt.begin(), cnt.end(), start);
  if (ids) {
    std::vector<int32_t> cur(cnt.begin(), cnt.end() - 1);
    for (int64_t i = 0; i < n; i++) ids[cur[(size_t)cells[i]]++] = (int32_t)i;  // ascending ids inside every cell
  }
  return MAVI_OK;
}

// the stencil of the GLOBAL periodic grid (mirrors for_each_neighbor, common.cuh)
int cell_neighbors(void *mm, int32_t cell, int32_t *out8, int32_t *n) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  const int R = m ? m->params.num_rows : 0, C = m ? m->params.num_cols : 0;
  if (!m || !out8 || !n || cell < 0 || cell >= R * C) return MAVI_ERR_BAD_PARAMS;
  const int col = cell / R, row = cell - col * R;
  int cnt = 0;
  for (int dc = -1; dc <= 1; dc++) {
    const int c2 = (col + dc + C) % C;
    for (int dr = -1; dr <= 1; dr++) {
      if (dc == 0 && dr == 0) continue;
      out8[cnt++] = c2 * R + (row + dr + R) % R;
    }
  }
  *n = cnt;
  return MAVI_OK;
}

int energies(void *mm, int32_t pe_mode, double *ke, double *pe) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m || !m->created) return MAVI_ERR_BAD_PARAMS;
  if (pe && pe_mode == 0) {
    m->set_error("n_gpus > 1: the exact O(N^2) potential energy is single-GPU only (SURVEY.md 8e); use pe_mode 1");
    return MAVI_ERR_UNSUPPORTED;
  }
  std::vector<double> k((size_t)m->G, 0.0), p((size_t)m->G, 0.0);
  int st = m->pool->run([&](int g) {
    return (int)m->api->energies(m->ranks[g].impl, pe_mode, ke ? &k[(size_t)g] : nullptr, pe ? &p[(size_t)g] : nullptr);
  });
  if (st) return m->first_error(st);
  if (ke) *ke = std::accumulate(k.begin(), k.end(), 0.0);
  if (pe) *pe = std::accumulate(p.begin(), p.end(), 0.0);
  return MAVI_OK;
}

int get_time(void *mm, int64_t *num_steps, double *time) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m) return MAVI_ERR_BAD_PARAMS;
  if (!m->created) {
    if (num_steps) *num_steps = m->num_steps0;
    if (time) *time = m->time0;
    return MAVI_OK;
  }
  return m->api->get_time(m->ranks[0].impl, num_steps, time);
}

int set_time(void *mm, int64_t num_steps, double time) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m) return MAVI_ERR_BAD_PARAMS;
  m->num_steps0 = num_steps;
  m->time0 = time;
  if (!m->created) return MAVI_OK;
  for (auto &r : m->ranks) m->api->set_time(r.impl, num_steps, time);
  return MAVI_OK;
}

int launch_count(void *mm, int64_t *n) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m || !n) return MAVI_ERR_BAD_PARAMS;
  *n = 0;
  if (!m->created) return MAVI_OK;
  for (auto &r : m->ranks) {
    int64_t v = 0;
    m->api->launch_count(r.impl, &v);
    *n += v;
  }
  return MAVI_OK;
}

int rebuild_count(void *mm, int64_t *n) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m || !n) return MAVI_ERR_BAD_PARAMS;
  *n = 0;
  if (!m->created) return MAVI_OK;
  for (auto &r : m->ranks) {
    int64_t v = 0;
    m->api->rebuild_count(r.impl, &v);
    *n += v;
  }
  return MAVI_OK;
}

int last_step_ms(void *mm, float *ms5) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m || !ms5) return MAVI_ERR_BAD_PARAMS;
  for (int i = 0; i < 5; i++) ms5[i] = 0.f;
  if (!m->created) return MAVI_OK;
  // per phase: the slowest device
  std::vector<float> all((size_t)m->G * 5, 0.f);
  int st = m->pool->run([&](int g) { return (int)m->api->last_step_ms(m->ranks[g].impl, &all[(size_t)g * 5]); });
  if (st) return m->first_error(st);
  for (int g = 0; g < m->G; g++)
    for (int i = 0; i < 5; i++) ms5[i] = std::max(ms5[i], all[(size_t)g * 5 + i]);
  return MAVI_OK;
}

int counters(void *mm, int64_t *out8) {
  Multi *m = reinterpret_cast<Multi *>(mm);
  if (!m || !out8) return MAVI_ERR_BAD_PARAMS;
  for (int i = 0; i < 8; i++) out8[i] = 0;
  if (!m->created) return MAVI_OK;
  std::vector<int64_t> all((size_t)m->G * 8, 0);
  int st = m->pool->run([&](int g) { return (int)m->api->counters(m->ranks[g].impl, &all[(size_t)g * 8]); });
  if (st) return m->first_error(st);
  for (int g = 0; g < m->G; g++)
    for (int i = 0; i < 8; i++) {
      const int64_t v = all[(size_t)g * 8 + i];
      out8[i] = (i == 0 || i == 6) ? std::max(out8[i], v) : out8[i] + v;  // steps / tile capacity are common, the rest adds up
    }
  return MAVI_OK;
}

}  // namespace mavi_multi
