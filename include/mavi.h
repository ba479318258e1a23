/* mavi.h — C ABI of libmavi_cuda.so, the B200 device backend for Mavi.jl's per-step hot path.
 *
 * The reference (Mavi.jl) has no FFI: its seam is Julia multiple dispatch on
 * IntCfg.device::DeviceMode (src/configs.jl:471-487) -> calc_forces!(system, chunks, device)
 * (src/integration.jl:112,159,197,226) and get_step_function (src/integration.jl:537-548,
 * src/rings/integration.jl:545).  A `CUDADevice <: DeviceMode` on the Julia side `ccall`s the
 * entry points below (binding shown in INTEGRATION.md).  Every entry point:
 *   - is extern "C", takes plain pointers/sizes, never retains host pointers after returning,
 *   - returns an int32 status (0 = MAVI_OK); mavi_last_error() gives the message.
 *
 * Memory contract: positions / velocities / forces are `Vector{SVector{2,T}}` on the Julia side
 * (src/states.jl:75-125) == contiguous T[2*N] (x0,y0,x1,y1,...).  All ids crossing the ABI are the
 * caller's ORIGINAL 0-based slot ids; the device reorders particles internally and un-permutes on
 * download.  Cell ids are 0-based linear ids  cell = (col-1)*num_rows + (row-1)  of the reference's
 * 1-based (row, col) (row 1 = TOP row, src/chunks.jl:129-130; row index fastest, src/chunks.jl:35-36).
 */
#ifndef MAVI_H
#define MAVI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAVI_ABI_VERSION 1
#define MAVI_MAX_SPACES 8

/* status codes (reference: Julia exceptions, see INTEGRATION.md "Error conventions") */
enum {
  MAVI_OK = 0,
  MAVI_ERR_BAD_PARAMS = 1,   /* error(...) in constructors, src/rings/rings.jl:233-262 */
  MAVI_ERR_OUT_OF_GRID = 2,  /* BoundsError from chunk_particles[...], src/chunks.jl:144-146 */
  MAVI_ERR_NAN = 3,
  MAVI_ERR_CUDA = 4,
  MAVI_ERR_NCCL = 5,
  MAVI_ERR_OUTSIDE_SPACE = 6, /* throw("Particles with ids=... outside space."), src/systems.jl:76-79 */
  MAVI_ERR_CAPACITY = 7,      /* halo / migration buffer overflow (multi-GPU only) */
  MAVI_ERR_UNSUPPORTED = 8
};

/* element type T of the state (src/states.jl:75-125; Float32 via NUM_T, src/init_states.jl:34,63).  The library holds
 * two builds of every kernel (real = double / float); MaviParams.dtype picks one per handle.  All `void *` state,
 * force, noise and Rings-info buffers are T[...]; parameters, times and energies are always double. */
enum { MAVI_F64 = 0, MAVI_F32 = 1 };

/* WallType subtypes, src/configs.jl:235-279 */
enum { MAVI_WALL_RIGID = 0, MAVI_WALL_PERIODIC = 1, MAVI_WALL_SLIPPERY = 2, MAVI_WALL_POTENTIAL = 3 };
/* GeometryCfg subtypes, src/configs.jl:28-152 */
enum { MAVI_GEOM_RECT = 0, MAVI_GEOM_CIRCLE = 1, MAVI_GEOM_LINES = 2 };
/* PotentialCfg used by PotentialWalls, src/configs.jl:341-397 */
enum { MAVI_POT_HARMTRUNC = 0, MAVI_POT_LJ = 1 };
/* PotentialWallMode, src/configs.jl:252-261 */
enum { MAVI_WALLMODE_OUTSIDE = 0, MAVI_WALLMODE_INSIDE = 1, MAVI_WALLMODE_REPULSION = 2 };
/* DynamicCfg subtypes, src/configs.jl:341-415, src/rings/configs.jl:95-107 */
enum { MAVI_DYN_LJ = 0, MAVI_DYN_HARMTRUNC = 1, MAVI_DYN_SZABO = 2, MAVI_DYN_RTP = 3, MAVI_DYN_RINGS = 4 };
/* stochastic terms: caller-supplied noise (exact parity) or device Philox4x32-10 (production) */
enum { MAVI_RNG_HOST_NOISE = 0, MAVI_RNG_PHILOX = 1 };

/* Line2D (src/configs.jl:95-117): normal/tangent/length are derived exactly as the ctor does. */
typedef struct MaviLine {
  double p1[2];
  double p2[2];
} MaviLine;

/* One (wall_type, geometry_cfg) pair of a SpaceCfg (src/configs.jl:285-302).  spaces[0] is the
 * main wall/geometry (get_main_wall / get_main_geometry, src/configs.jl:304-312). */
#define MAVI_MAX_POT_TYPES 4
typedef struct MaviSpace {
  int32_t wall;          /* MAVI_WALL_* */
  int32_t geom;          /* MAVI_GEOM_* */
  double rect_bl[2];     /* RectangleCfg.bottom_left */
  double rect_len;       /* RectangleCfg.length */
  double rect_h;         /* RectangleCfg.height */
  double circ_center[2]; /* CircleCfg.center */
  double circ_radius;    /* CircleCfg.radius */
  const MaviLine *lines; /* LinesCfg.lines (copied at create) */
  int32_t n_lines;
  int32_t pot_kind;      /* PotentialWalls.potential: MAVI_POT_* */
  double pot[4];         /* HarmTrunc: k_rep,k_atr,dist_eq,dist_max; LJ: sigma,epsilon */
  int32_t pot_mode;      /* MAVI_WALLMODE_* */
  int32_t n_pot_types;   /* 0: `pot` for every particle; > 0: PotentialVector (src/configs.jl:454-463), pot_types[t] for particles
                          * of type t + 1 (get_particle_type: the ring type — Mavi.Rings states only), all of kind pot_kind */
  double pot_types[MAVI_MAX_POT_TYPES][4];
} MaviSpace;

/* RingsCfg + RingsState layout (src/rings/configs.jl:95-107, src/rings/states.jl:74-124).
 * Scalar idx of particle p (0-based) of ring r (0-based) = r*n_max + p (src/rings/states.jl:137-139). */
typedef struct MaviRingsParams {
  int32_t num_types;
  int32_t n_max;             /* num_max_particles = size(rings_pos, 1) */
  int64_t num_rings;
  const double *p0, *relax_time, *vo, *mobility, *rot_diff, *k_area, *k_spring, *l_spring; /* [num_types] */
  const int32_t *num_particles; /* [num_types] */
  const double *interaction;    /* [num_types][num_types][4] = k_rep,k_atr,dist_eq,dist_max (InteractionMatrix, src/rings/configs.jl:59-89) */
  const int32_t *types;         /* [num_rings], 1-based ring types, or NULL when the state has no types */
} MaviRingsParams;

typedef struct MaviParams {
  uint32_t struct_size; /* sizeof(MaviParams), ABI check */
  int32_t dtype;        /* MAVI_F64 (default) / MAVI_F32 */
  int64_t n;            /* number of particle slots = length(state.pos) */

  int32_t n_spaces;
  int32_t _pad0;
  MaviSpace spaces[MAVI_MAX_SPACES];

  /* Chunks (src/chunks.jl:26-40): grid over the bounding box of the geometry (src/systems.jl:14-28).
   * num_cols == 0 -> chunks === nothing -> all-pairs path (src/integration.jl:197-224). */
  double grid_bl[2];
  double grid_len;
  double grid_h;
  int32_t num_cols;
  int32_t num_rows;

  int32_t dynamics; /* MAVI_DYN_* */
  int32_t _pad1;
  /* LJ: sigma,epsilon | HarmTrunc: k_rep,k_atr,dist_eq,dist_max |
   * Szabo: vo,mobility,relax_time,k_rep,k_adh,r_eq,r_max,rot_diff | RTP: vo,sigma,epsilon,tumble_rate */
  double dyn[8];
  double particle_radius; /* particle_radius(dynamic_cfg) as computed by the host (src/configs.jl:418-421); Rings: minimum over types */
  const MaviRingsParams *rings; /* MAVI_DYN_RINGS only */

  double dt; /* IntCfg.dt */

  int32_t rng_mode; /* MAVI_RNG_* */
  int32_t n_gpus;   /* 0 / 1: one GPU (MaviParams.device).  G > 1: ONE process drives G GPUs (devices device .. device+G-1)
                     * through this one handle — the reference's driver is a single process (src/run_system.jl:7-23,
                     * src/systems.jl:73-114): the library partitions the state into x-slabs of cell columns on
                     * mavi_upload_state, steps all slabs concurrently (halo / migration over NCCL, NVLink) and un-permutes
                     * on download.  Needs a single periodic rectangle with chunks; not Mavi.Rings; stream must be NULL. */
  uint64_t seed;

  int32_t device; /* CUDA device ordinal */
  int32_t flags;  /* MAVI_FLAG_* */
  void *stream;   /* cudaStream_t to enqueue on, or NULL for the legacy default stream */

  /* x-slab domain decomposition with ONE PROCESS PER GPU (torchrun-style launchers; exclusive with n_gpus > 1).
   * world<=1 -> not used. */
  int32_t rank;
  int32_t world;
  const void *nccl_unique_id; /* ncclUniqueId bytes (128), same on every rank */
  int64_t n_global;           /* total particle count over all ranks (0 -> n) */
} MaviParams;

/* flags */
#define MAVI_FLAG_RESORT_EVERY_STEP 1 /* full update_chunks! rebuild every step instead of the incremental tile repair (A/B testing) */
#define MAVI_FLAG_NO_FORCE_CARRY 4    /* Newton steps: always run the full first force pass instead of carrying F2 / the drift over from the previous step (A/B testing; results are bit-identical) */
#define MAVI_FLAG_SMALL_BLOCKS 8      /* testing: 3 tile columns per CTA, so that block splitting, chunking and the slab overlap path run on small systems */
#define MAVI_FLAG_SLAB_SELF 16        /* world == 1 only: run the x-slab machinery (halo columns, emigrant records, two-stream step pipeline) with this rank as its own periodic neighbour, device copies instead of NCCL; state moves through mavi_upload_local / mavi_download_local.  Lets the multi-GPU path be tested and profiled on one GPU */
#define MAVI_FLAG_LEGACY_STAGING 32    /* A/B testing: the round-1 force kernels (one CTA per tile block, cp.async staging behind CTA barriers) instead of the persistent producer/consumer kernels with bulk-async staging; results are bit-identical */
#define MAVI_FLAG_TIGHT_TILES 2       /* testing: tile capacity without slack, so that the overflow -> rebuild -> resume path is exercised */

typedef struct MaviHandle MaviHandle;

/* ---- lifetime -------------------------------------------------------------------------------
 * replaces: System(...) ctor, src/systems.jl:73-114 (force buffers, Chunks build, first update_chunks!)
 *           RingsSystem(...) ctor, src/rings/rings.jl:231-291 */
int32_t mavi_create(const MaviParams *params, MaviHandle **out);
int32_t mavi_destroy(MaviHandle *h);
int32_t mavi_abi_version(void);

/* ---- state movement (the only host<->device copies) -------------------------------------------
 * replaces direct reads/writes of system.state.{pos,vel,pol_angle,pol}.
 * `second` is vel (SecondLawState, T[2n]) or pol_angle (SelfPropelledState, T[n]) or pol
 * (RingsState, T[num_rings]).  active_mask (n bytes, may be NULL = all active) is the ParticleIds mask
 * (src/states.jl:27-52).  The constructor-time inside check (src/space_checks.jl:9-37) runs here. */
int32_t mavi_upload_state(MaviHandle *h, const void *pos, const void *second, const uint8_t *active_mask, int64_t n);
int32_t mavi_download_state(MaviHandle *h, void *pos, void *second);
/* get_forces(system), src/systems.jl:117 */
int32_t mavi_download_forces(MaviHandle *h, void *forces);
/* multi-GPU only: original ids of the particles this rank currently owns, and their count */
int32_t mavi_local_count(MaviHandle *h, int64_t *n_local);
int32_t mavi_download_local(MaviHandle *h, int64_t *ids, void *pos, void *second, void *forces);
/* multi-GPU x-slabs (MaviParams.world > 1, one process per GPU; the reference has no distributed path — it splits the
 * same loop over cell columns across threads, src/integration.jl:159-194).  Rank 0 makes the NCCL id, the host
 * broadcasts its 128 bytes to all ranks (MaviParams.nccl_unique_id); every rank uploads the particles whose cell
 * column it owns (columns are split contiguously, the first num_cols % world ranks get one extra) with global ids. */
int32_t mavi_nccl_unique_id(void *out128);
int32_t mavi_upload_local(MaviHandle *h, const int64_t *ids, const void *pos, const void *second, int64_t n_local);

/* ---- the hot path ---------------------------------------------------------------------------
 * mavi_step: nsteps x (newton_step! | szabo_step! | rtp_step!  src/integration.jl:507-535 |
 *                      Rings step!  src/rings/integration.jl:522-543), chosen like get_step_function.
 * host_noise (MAVI_RNG_HOST_NOISE): per step  Szabo T[n] (randn, src/integration.jl:460) |
 *   RTP T[2n] (u, u2 pairs, src/integration.jl:493-495) | Rings T[num_rings] (randn, src/rings/integration.jl:348);
 *   may be NULL when the stochastic amplitude is zero or for Newton dynamics.  The row is indexed by the ORIGINAL
 *   particle id: in x-slab mode n is MaviParams.n_global and every rank passes the same full global rows. */
int32_t mavi_step(MaviHandle *h, int64_t nsteps, const void *host_noise);
/* clean_forces! + update_chunks! + calc_forces! (+ Rings: springs, area forces) + calc_walls_forces!:
 * the force state a reference system holds right after these calls (src/integration.jl:508-511) */
int32_t mavi_calc_forces(MaviHandle *h);
/* update_chunks!(system.chunks), src/chunks.jl:150-163, src/integration.jl:54-59 */
int32_t mavi_bin(MaviHandle *h);
/* cell_of_particle[n] (0-based linear cell id, -1 for inactive), counts[num_rows*num_cols]
 * (== num_particles_in_chunk, row fastest).  Either pointer may be NULL. */
int32_t mavi_download_cells(MaviHandle *h, int32_t *cell_of_particle, int32_t *counts);
/* chunk_particles as CSR: start[num_cells+1], ids[n_active] ascending ids inside each cell
 * (== the reference's fill order, src/chunks.jl:153-155).  Either pointer may be NULL. */
int32_t mavi_download_cell_lists(MaviHandle *h, int32_t *start, int32_t *ids);
/* neighbour stencil actually used by the pair kernels for cell `cell`: up to 8 neighbour cell ids
 * (the reference's half stencil src/chunks.jl:61-118 united with its mirror image); returns count in *n. */
int32_t mavi_cell_neighbors(MaviHandle *h, int32_t cell, int32_t *out8, int32_t *n);

/* update_particle_chunk! (src/chunks.jl:120-147) evaluated on the HOST with the device's exact arithmetic (Base.div as trunc
 * of the real quotient, the clamp of index n+1): cell_out[i] = 0-based linear cell id of point i, -1 outside the grid.
 * pos is T[2n] with T = params->dtype; only the grid fields of params are read.  Needs no GPU: host-side callers that
 * route particles to slabs (one process per GPU) use it so that host and device can never disagree on a cell. */
int32_t mavi_cells_of_points(const MaviParams *params, const void *pos, int64_t n, int32_t *cell_out);

/* ---- quantities, src/quantities.jl ----------------------------------------------------------
 * ke = kinetic_energy (:12-18; NaN for states without vel).  pe_mode 0: exact all-pairs LJ
 * potential_energy (:46-66, O(N^2)); 1: sum over the cell-stencil pair set (labelled deviation). */
int32_t mavi_energies(MaviHandle *h, int32_t pe_mode, double *ke, double *pe);

/* ---- Rings info, src/rings/rings.jl:118-217 -------------------------------------------------
 * areas[num_rings], cms[2*num_rings], cont_pos[2*n] (continuos_pos).  Any pointer may be NULL. */
int32_t mavi_rings_download_info(MaviHandle *h, void *areas, void *cms, void *cont_pos);

/* ---- particle contact lists, src/rings/neighbors.jl:11-136 (RingsSystem p_neighbors_cfg, src/rings/rings.jl:143-158) ----
 * mavi_rings_set_neighbors: NeighborsCfg(only_count, type, tol).  mode MAVI_NEIGH_COUNT == only_count=true,
 * MAVI_NEIGH_LIST keeps up to MAVI_NEIGH_MAX = 15 neighbour ids per particle (num_max_neighbors, src/rings/rings.jl:145);
 * type_all != 0 is type = :all, 0 is :rings (particles of the same ring are not neighbours).  Call it after
 * mavi_create and before mavi_upload_state to have the constructor's first forces! fill the lists like the reference
 * (src/rings/rings.jl:280-288); every later forces! (mavi_step, mavi_calc_forces) cleans and refills them
 * (neigh_clean!, src/rings/integration.jl:362; neigh_update!, :42).
 * mavi_rings_download_neighbors: count[n] (get_neigh_count) and list[n][15] (get_neigh_list; 0-based particle ids in
 * ASCENDING order, -1 padded — the reference appends in pair-enumeration order and compares sorted lists).  Either
 * pointer may be NULL.  MAVI_ERR_CAPACITY when a particle has more than 15 contacts in list mode (the reference's
 * BoundsError). */
enum { MAVI_NEIGH_OFF = 0, MAVI_NEIGH_COUNT = 1, MAVI_NEIGH_LIST = 2 };
#define MAVI_NEIGH_MAX 15
int32_t mavi_rings_set_neighbors(MaviHandle *h, int32_t mode, int32_t type_all, double tol);
int32_t mavi_rings_download_neighbors(MaviHandle *h, int32_t *count, int32_t *list);

/* ---- sources, sinks and a variable number of rings, src/rings/sources.jl, src/rings/states.jl:173-227 -------------------
 * RingsSystem(source_cfg = [SourceCfg(...), SinkCfg(...), ...]) with RingsState(active_state = ActiveState(mask)).  At the head
 * of every step! (src/rings/integration.jl:353-358,523-526), after update_cms!, the list is processed in order:
 *   sink   : every active ring whose centre of mass (info.cms, as of this step's update_cms!) lies inside the geometry is
 *            removed (remove_ring!);
 *   source : size[0] x size[1] spawn areas (bounding box of spawn_pos + pad, laid out from bottom_left with `offset`); every
 *            area that holds no active particle (is_inside(pos, bbox, pad)) spawns a ring into the FIRST free ring slot
 *            (add_ring!: rings_pos[:, slot] = spawn_pos + shift, pol = spawn_pol, uid = max(uids) + 1, cms primed);
 * then update_ids! recomputes the active ids.  num_spawn_pos must equal n_max.
 * ring_active: [num_rings] ActiveState mask (non-zero = active) of the uploaded state; NULL = all active (FixRingsIds).
 * spawn_draws: the rand(rng) values consumed by `spawn_pol = :random` sources (pol = draw * 2 pi), in spawn order — the
 *   reference draws them from system.rng, which no device stream reproduces; NULL = Philox4x32 keyed (seed, spawn count).
 * Call after mavi_create and before mavi_upload_state. */
enum { MAVI_SRC_SOURCE = 0, MAVI_SRC_SINK = 1 };
typedef struct MaviSourceSink {
  int32_t kind;            /* MAVI_SRC_* */
  int32_t num_spawn_pos;   /* source: length(SourceCfg.spawn_pos) */
  const double *spawn_pos; /* source: [2 * num_spawn_pos] */
  double bottom_left[2];   /* source: SourceCfg.bottom_left */
  double spawn_pol;        /* source: SourceCfg.spawn_pol; NaN = :random */
  double pad;              /* source */
  double offset[2];        /* source */
  int32_t size[2];         /* source */
  int32_t sink_geom;       /* sink: MAVI_GEOM_RECT / MAVI_GEOM_CIRCLE */
  int32_t _pad;
  double sink_rect_bl[2], sink_rect_len, sink_rect_h; /* sink: RectangleCfg */
  double sink_circ_center[2], sink_circ_radius;       /* sink: CircleCfg */
} MaviSourceSink;
int32_t mavi_rings_set_sources(MaviHandle *h, const MaviSourceSink *list, int32_t n, const uint8_t *ring_active,
                               const double *spawn_draws, int64_t n_draws);
/* VarRingsIds after the last step: mask[num_rings], uids[num_rings] (either may be NULL), number of active rings */
int32_t mavi_rings_download_active(MaviHandle *h, uint8_t *ring_active, int64_t *uids, int64_t *num_active);

/* ---- ring invasions, src/rings/integration.jl:379-520 ----------------------------------------------------------------
 * RingsIntCfg(invasions_cfg = InvasionsCfg(steps_to_update), r_chunks_cfg = ChunksCfg(r_cols, r_rows)).  Every steps_to_update
 * steps (update_invasions!, :509-520; before the forces of that step) the rings are binned by centre of mass into the ring
 * chunks (update_chunks!(r_chunks), :18-23) and, for every pair of rings in the same or adjacent ring chunks, the particles of
 * one that lie inside the polygon of the other (polygons_intersect: ray casting with point_line_intersect on ring_points) are
 * listed.  r_cols = r_rows = 0: no ring chunks, every pair of rings is tested (check_invasions!(system, ::Nothing)).
 * mavi_rings_download_invasions: info.invasions.list of the last check as triples (invasor ring, invaded ring, scalar particle
 * id), 0-based, sorted lexicographically (the reference lists them in pair-enumeration order); *n = their number. */
int32_t mavi_rings_set_invasions(MaviHandle *h, int32_t steps_to_update, int32_t r_cols, int32_t r_rows);
int32_t mavi_rings_download_invasions(MaviHandle *h, int64_t *n, int32_t *triples, int64_t cap);

/* TimeInfo, src/systems.jl:30-33 (time += dt accumulated in Float64, src/integration.jl:500-503) */
int32_t mavi_get_time(MaviHandle *h, int64_t *num_steps, double *time);
int32_t mavi_set_time(MaviHandle *h, int64_t num_steps, double time);

/* blocks until all enqueued work is done and returns the deferred device error word as status */
int32_t mavi_sync(MaviHandle *h);
int32_t mavi_last_error(MaviHandle *h, char *buf, int32_t n);

/* ---- instrumentation (bench.py / tests) ------------------------------------------------------ */
/* number of kernels this handle has launched since creation */
int32_t mavi_launch_count(MaviHandle *h, int64_t *n);
/* number of tile-overflow -> full rebuild events since creation */
int32_t mavi_rebuild_count(MaviHandle *h, int64_t *n);
/* device time of the last mavi_step call per phase, CUDA events on the launching stream:
 * ms[0]=bin+sort ms[1]=pass A ms[2]=pass B (or the single fused pass) ms[3]=exchange ms[4]=total */
int32_t mavi_last_step_ms(MaviHandle *h, float *ms5);
int32_t mavi_set_profiling(MaviHandle *h, int32_t on);
/* totals since the last upload: out8[0] steps run, [1] particles re-binned (cell changes = what update_chunks! would move,
 * src/chunks.jl:150-163), [2] inter-tile movers, [3] tiles repaired, [4] slab emigrants, [5] tile-overflow rebuilds,
 * [6] tile capacity (slots), [7] number of tiles.  Synchronises. */
int32_t mavi_counters(MaviHandle *h, int64_t *out8);

#ifdef __cplusplus
}
#endif
#endif /* MAVI_H */
