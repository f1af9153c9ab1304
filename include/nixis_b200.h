/*
 * nixis_b200.h -- C-ABI of libnixis_b200.so (hand-written sm_100a CUDA).
 *
 * The reference (MightyBOBcnc/nixis) has no FFI layer: its hot path is plain
 * Python functions that nixis.py imports by name.  This header is therefore the
 * boundary a maintainer binds with ctypes (INTEGRATION.md shows the stub); every
 * entry point cites the reference function it replaces.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch / C++ types.
 *  - Pointers are DEVICE pointers unless the name ends in _h / _host.
 *  - Every function returns 0 on success or a negative nxb_status; it never
 *    throws.  nxb_last_error() returns the text of the last failure (thread-local).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Calls are asynchronous on that stream unless stated otherwise.
 *  - Thread-compatible, not thread-safe per handle.
 *  - Device state is FP32 (heights, water, sediment, unit-sphere positions);
 *    integer tables (perm, cells, adjacency) are bit-exact w.r.t. the reference.
 */
#ifndef NIXIS_B200_H
#define NIXIS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NXB_ABI_VERSION 1

typedef enum {
    NXB_OK = 0,
    NXB_ERR_CUDA = -1,        /* a CUDA runtime call failed (see nxb_last_error) */
    NXB_ERR_ARG = -2,         /* bad argument */
    NXB_ERR_OVERFLOW = -3,    /* adjacency row overflow (inconsistent winding) */
    NXB_ERR_UNSUPPORTED = -4
} nxb_status;

/* positions on the device: unit-sphere xyz in .x .y .z, .w unused (16-byte loads) */
typedef struct { float x, y, z, w; } nxb_float4;

/* ---- library ----------------------------------------------------------- */
int nxb_version(void);
int nxb_last_error(char *buf, int len);
/* sm count, clock (kHz), total memory of the CURRENT device */
int nxb_device_info(int *sm_count, int *sm_clock_khz, int64_t *mem_bytes, int *cc_major, int *cc_minor);
/* FP32 FFMA microbenchmark on the current device: dependent-free FFMA chains on
 * every SM; returns achieved TFLOP/s (FMA = 2 flops).  Synchronous.  Used as the
 * measured FP32 roofline denominator (MEASURED_PEAKS.json has no FP32 entry). */
int nxb_ffma_peak(int iters, double *tflops_out);

/* ---- opensimplex.py ---------------------------------------------------- */
/* opensimplex.py:90-112 init(seed) incl. the int32 truncation of `over`.  HOST. */
int nxb_init_perm(int64_t seed, int32_t *perm_host, int32_t *pgi_host);
/* Pack perm / perm_grad_index_3D into the device lookup tables the kernels stage
 * in shared memory.  Synchronous (tiny H2D).  handle_out receives an opaque pointer. */
int nxb_tables_create(const int32_t *perm_host, const int32_t *pgi_host, void **handle_out);
int nxb_tables_destroy(void *handle);
/* opensimplex.py:257-263 noisearr3d / :144-150 noisearr2d / :762-768 noisearr4d:
 * one evaluation per element, FP32 in/out. */
int nxb_noise3_f32(void *tables, const float *x, const float *y, const float *z, int64_t n, float *out, void *stream);
int nxb_noise2_f32(void *tables, const float *x, const float *y, int64_t n, float *out, void *stream);
int nxb_noise4_f32(void *tables, const float *x, const float *y, const float *z, const float *w, int64_t n, float *out, void *stream);

/* ---- terrain.py -------------------------------------------------------- */
/* terrain.py:12-59 sample_noise + sample_octaves fused over all octaves:
 *   out[v] = (init ? init[v] : 0) + sum_o ((noise3(xyz[v]*freq[o]) + 1) * 0.5) * amp[o]
 * freq_host/amp_host: n_oct doubles each, already advanced octave by octave the way
 * terrain.py:44-45 does (freq *= roughness; amp *= persistence), in UNIT-sphere units
 * (the reference's n_freq/world_radius applied to radius-scaled verts is the same number).
 * minmax (nullable, device float[2]) receives min/max of out over [0,n) merged into its
 * current content (initialise with +inf/-inf), for terrain.py:50-53 and util.rescale. */
int nxb_fbm3_f32(void *tables, const nxb_float4 *xyz_unit, int64_t n, int n_oct,
                 const double *freq_host, const double *amp_host,
                 const float *init, float *out, float *minmax, void *stream);
/* The same with the reference's own float64 vertices (device double[n][3], multiplied by `scale`
 * first: nixis.py:249 `points *= world_radius`; pass 1.0 for radius-scaled vertices).
 * nr_host[o] = n_freq_o / world_radius (terrain.py:43), amp_host[o] = octave amplitude in output units.
 * Lattice coordinates, cell and candidate selection are float64 in the reference's operation order, so
 * the evaluated lattice points are exactly the reference's (ties included); the contributions and
 * the accumulation are FP32.  nxb_fbm3_f32 is this function on float4 positions promoted to double. */
int nxb_fbm3_pos64_f32(void *tables, const double *verts, double scale, int64_t n, int n_oct,
                       const double *nr_host, const double *amp_host,
                       const float *init, float *out, float *minmax, void *stream);
/* Reference-exact mode of the same function: IEEE double, no FMA contraction, the reference's
 * operation order -> bit-identical to the numba output for the same float64 vertices.
 * verts: double[n][3] (device), multiplied by `scale` first (nixis.py:249 `points *= world_radius`;
 * pass 1.0 for vertices that are already radius-scaled); nr_host[o] = n_freq_o / world_radius,
 * ns_host[o] = n_amp_o / world_radius (terrain.py:43); out = init + sum_o (e+1)*0.5*ns*radius. */
int nxb_fbm3_f64(void *tables, const double *verts, int64_t n, int n_oct,
                 const double *nr_host, const double *ns_host, double radius, double scale,
                 const double *init, double *out, void *stream);
/* 4-D variant (no reference driver exists; w coordinate = w_host[o] per octave). */
int nxb_fbm4_f32(void *tables, const nxb_float4 *xyz_unit, int64_t n, int n_oct,
                 const double *freq_host, const double *amp_host, const double *w_host,
                 const float *init, float *out, float *minmax, void *stream);
/* terrain.py:61-72 make_bool_elevation_mask */
int nxb_mask_le_f32(const float *h, int64_t n, float level, uint8_t *mask, void *stream);

/* ---- util.py: mesh ------------------------------------------------------ */
/* util.py:17-50 create_mesh -> meshzoo.icosa_sphere(k) (layout: SURVEY App. B,
 * parity unpinned).  Vertices [v_begin, v_end) of the k-division icosphere, closed form.
 * xyz_f32 and/or xyz_f64 (double[.][3]) may be NULL. */
int nxb_mesh_icosa_points(int k, int64_t v_begin, int64_t v_end, nxb_float4 *xyz_f32, double *xyz_f64, void *stream);
/* triangles [t_begin, t_end) as int32[.][3] */
int nxb_mesh_icosa_cells(int k, int64_t t_begin, int64_t t_end, int32_t *cells, void *stream);
/* util.py:591-662 build_adjacency + sort_adjacency for rows [v_begin, v_end) of the closed-form icosphere
 * WITHOUT the cell array or any whole-mesh table (a multi-GPU rank builds only its own rows): the
 * closed-form triangle generator is scanned, corners inside the range are collected, each vertex
 * sorts its <= 6 incident triangles by index (= the reference's append order) and walks its ring.
 * adj_sorted int32[n][6] (global ids, -1 pad), adj_unsorted nullable (the build_adjacency rows);
 * workspace: nxb_mesh_icosa_adj_rows_workspace(n) bytes (52 B per row).  Synchronous at the end. */
int64_t nxb_mesh_icosa_adj_rows_workspace(int64_t n_rows);
int nxb_mesh_icosa_adj_rows(int k, int64_t v_begin, int64_t v_end, int32_t *adj_sorted, int32_t *adj_unsorted,
                            void *workspace, void *stream);
/* double[n][3] * scale -> float4 (used to ingest caller-supplied float64 vertices) */
int nxb_xyz_f64_to_f32(const double *xyz_f64, int64_t n, double scale, nxb_float4 *xyz_f32, void *stream);

/* ---- util.py: equirectangular export --------------------------------------- */
/* util.py:290-308 make_ll_arr: xyz (double[height][width][3]) of every pixel's lat/lon */
int nxb_ll_grid_f64(int width, int height, double radius, double *xyz_out, void *stream);
/* The 3 nearest vertices of the k-division icosphere (radius-scaled) to each query position --
 * what the reference gets from scipy KDTree(points).query(ll, k=3) (nixis.py:270-283), found
 * analytically from the closed-form mesh.  dists double[n][3] ascending, ids int64[n][3]. */
int nxb_ico_nearest3_f64(int k, double radius, const double *query_xyz, int64_t n, double *dists, int64_t *ids, void *stream);
/* util.py:343-367 make_gray_array: inverse-distance blend, float64, reference operation order,
 * int() truncation.  colors: double[V]; out int32[n]. */
int nxb_idw_gray_f64(const double *dists, const int64_t *ids, const double *colors, int64_t n, int32_t *out, void *stream);
/* One export map per launch, straight from a device-resident field (nixis.py:349, 386-389, 417 ->
 * util.py:393-429): colour = rescale(field, lower, upper) + add over the given [x_min, x_max]
 * (util.py:143), truncated to uint16 first when quantize_u16 (".astype('uint16')", nixis.py:386,389),
 * blended like make_gray_array and stored as uint8 / uint16 (out_bits).  field_kind 0 = float32[V],
 * 1 = uint8[V] (masks). */
int nxb_idw_map(const double *dists, const int64_t *ids, const void *field, int field_kind, int64_t n,
                double x_min, double x_max, double lower, double upper, double add, int quantize_u16,
                int out_bits, void *out, void *stream);

/* ---- per-vertex climate kernels ---------------------------------------------------
 * verts: device float64 [n][3] (the reference's `points`, already scaled by radius).  Every driver
 * of the reference (360 rotations, 360 days) is ONE pass over the vertices. */
/* climate.py:345-372 assign_surface_temp (the altitude term is multiplied by the literal 0 there) */
int nxb_climate_surface_temp_f32(const double *verts, int64_t n, double radius, double tilt, float *out, void *stream);
/* climate.py:415-448 sample_insolation, repeated for rot_deg[0..n_rot) (HOST, degrees) the way
 * brute_daily_insolation (:450-490) and calc_insolation_slice (:503-536) drive it, for each of
 * tilt_deg[0..n_tilt) (HOST, degrees): arr float32 [n_tilt][n] += ..., float32 rounding per rotation.
 * tilt_scratch: device double[2*n_tilt].  n_rot, n_tilt <= 360. */
int nxb_climate_insolation_f32(const double *verts, int64_t n, double radius, const double *rot_deg, int n_rot,
                               const double *tilt_deg, int n_tilt, double *tilt_scratch, float *arr, void *stream);
/* the 181 lookup vertices of calc_insolation_slice (climate.py:507-515, util.py:80-88), HOST output */
int nxb_climate_slice_verts(double radius, double *verts_host);
/* climate.py:193-201 calculate_seasonal_tilt (host) */
double nxb_climate_seasonal_tilt(double axial_tilt, double degrees);
/* climate.py:551-577 interpolate_insolation against n_tab tables float32 [n_tab][181]; n_tab > 1 sums
 * the float32 daily values in table order (calc_yearly_insolation, :579-597) */
int nxb_climate_interpolate_f32(const double *verts, int64_t n, double radius, const float *tables, int n_tab,
                                float *out, void *stream);

/* ---- util.py: adjacency -------------------------------------------------- */
/* util.py:591-613 build_adjacency.  cells int32[T][3]; adj int32[V][6], -1 padded.
 * workspace: device scratch of nxb_adj_build_workspace(V) bytes.  Deterministic
 * (slot = rank of the triangle index among the triangles around the vertex).
 * Synchronous at the end (reads back the overflow flag). */
int64_t nxb_adj_build_workspace(int64_t V);
int nxb_adj_build(const int32_t *cells, int64_t T, int64_t V, int32_t *adj, void *workspace, void *stream);
/* util.py:623-662 sort_adjacency, race-free: reads adj_in, writes adj_out (must differ). */
int nxb_adj_sort(const int32_t *adj_in, int32_t *adj_out, int64_t V, void *stream);

/* ---- util.py: rescale ---------------------------------------------------- */
/* min/max of x merged into minmax[2] (device; initialise with +inf/-inf via nxb_minmax_reset) */
int nxb_minmax_reset(float *minmax, void *stream);
int nxb_minmax_f32(const float *x, int64_t n, float *minmax, void *stream);
/* util.py:110-175 rescale.  x_min/x_max are passed by the caller (after the u_min/u_max
 * widening and, multi-GPU, the all-reduce).  mode: 0 None, 1 'lower', 2 'upper'.
 * out may alias x. */
int nxb_rescale_f32(const float *x, int64_t n, float x_min, float x_max, float lower, float upper,
                    int has_mid, float mid, int mode, float *out, void *stream);
/* util.py:178-254 power_rescale, pass 1: statistics of the selected elements
 * (sel_mode 1 -> mask true, 0 -> mask false) in ONE ordered pass.  summary4 (device float[4])
 * receives (has, F, U, M): has = 1 if anything is selected, F = first selected value,
 * M = min selected, U = max over selected values that are >= the min of the selected values
 * before them.  The reference's sequential if/elif scan (util.py:203-214) is then
 *     mask_lower = min(x_max, M);  mask_upper = max(x_min, U, F >= x_max ? F : -inf)
 * Shards combine in index order: F = first shard's F, M = min, U = max(U_a, U_b, F_b >= M_a ? F_b : -inf). */
int nxb_power_summary_f32(const float *x, const uint8_t *mask, int64_t n, int sel_mode,
                          float *summary4, void *stream);
/* Pass 2: out = selected ? pow((x-lo)/(hi-lo), power)*(hi-lo)+lo : x.  out may alias x.
 * sel_mode -1 = mode None (nothing selected).
 * `shift` is subtracted from every element afterwards (nixis.py:359 `height -= ocean_level`
 * fused; pass 0 otherwise). */
int nxb_power_apply_f32(const float *x, const uint8_t *mask, int64_t n, int sel_mode,
                        float lo, float hi, float power, float shift, float *out, void *stream);

/* ---- erosion.py ---------------------------------------------------------- */
/* erosion.py:34-40 calc_distance for every adjacency slot, once, in FP64, stored as FP32:
 * dist[v*6+q] = |nodes[v] - nodes[adj[v][q]]| (0 for -1 pads).  The reference measures distances on
 * the undisplaced sphere positions (erosion.py:227-229), so they are constant over the run. */
int nxb_edge_lengths_f64(const double *nodes /*[.][3]*/, const int32_t *adj, int64_t n_own, float *dist, void *stream);
/* same for the closed-form icosphere, rows [v_begin, v_end) with GLOBAL vertex ids in adj_rows */
int nxb_mesh_icosa_edge_lengths(int k, const int32_t *adj_rows, int64_t v_begin, int64_t v_end, double radius,
                                float *dist, void *stream);
/* Tile plan of the sweep (see csrc/nxb_erosion_plan.cuh): per 256-vertex tile the contiguous halo
 * segments to stage in shared memory, the adjacency re-encoded as 16-bit tile-local codes, and the
 * per-tile constants of the implicit-adjacency kinds.
 * adj: int32[n_own][6] with indices in [0, capacity) (own vertices first, then halo slots of a
 * multi-GPU shard).  capacity = allocated ELEMENTS of every h/w/s buffer later passed to the step
 * (multiple of 256, >= round_up(n_own, 256)).  plan_mem: nxb_erode_plan_bytes(n_own) bytes, 16-byte
 * aligned.  stats_host (nullable) int32[6]: tiles, irregular tiles, max halo slots, affine tiles
 * (implicit adjacency: the sweep reads no adjacency codes for them), affine tiles that also qualify
 * for one stored length per edge, two-piece tiles (a mesh-row end inside the tile: two sets of
 * constants + 4 exception vertices).  Synchronous. */
int64_t nxb_erode_plan_bytes(int64_t n_own);
int nxb_erode_plan_build(const int32_t *adj, int64_t n_own, int64_t capacity, void *plan_mem,
                         int32_t *stats_host, void *stream);
/* One stored length per edge: dist3[v][i] = dist[v][q] of v's i-th larger-numbered neighbour (slot
 * order), nxb_erode_dist3_floats(n_own) floats.  With it the sweep streams 12 B/vertex of edge
 * lengths instead of 24 on the tiles the plan marks (36 B per vertex-iteration instead of 48).
 * Pass dist3 = NULL to the sweep functions to stream the full table everywhere. */
int64_t nxb_erode_dist3_floats(int64_t n_own);
int nxb_erode_dist3_build(const int32_t *adj, const float *dist, int64_t n_own, float *dist3, void *stream);
/* erosion.py:197-279 erosion_iteration3 for vertices [0, n_own), FP32 state, ping-pong buffers
 * (reads *_in, writes *_out; no copy-back pass).  Heights and water live INTERLEAVED, hw = float[capacity][2]
 * = {height, water} per vertex: a neighbour's height and water are always read together, so a run of
 * neighbours is one bulk copy and a boundary value is one 8-byte peer store; sediment is a separate
 * float[capacity].  `rain` is added to every water value read (erosion.py:182-183 `water += rain_amount`
 * fused).  dist: float[round_up(n_own,256)*6]. */
int nxb_erode3_plan_step_f32(const void *plan_mem, const int32_t *adj, const float *dist, const float *dist3,
                             const float *hw_in, const float *s_in, float *hw_out, float *s_out,
                             int64_t n_own, float rain, void *stream);
/* erosion.py:180-184: the erode_terrain3 loop, n_sweeps sweeps issued from C (one launch each, chained
 * with programmatic dependent launch).  Sweep 0 reads buffer set A and writes B, sweep 1 reads B ...:
 * the result is in A when n_sweeps is even, in B when it is odd. */
int nxb_erode3_run_f32(const void *plan_mem, const int32_t *adj, const float *dist, const float *dist3,
                       float *hw_a, float *s_a, float *hw_b, float *s_b,
                       int64_t n_own, float rain, int64_t n_sweeps, void *stream);
/* The same loop on one shard of a multi-GPU run, each sweep fused with the halo exchange in ONE
 * kernel: boundary {height, water} pairs are stored straight into the peers' halo slots over NVLink as
 * they are computed and the last CTA raises this rank's flag in every peer; a one-warp kernel in front
 * of each sweep waits for the peers' flags of the previous one.  The plan's tile descriptors carry each
 * tile's range of send_entries (device array of {int32 dst, uint16 vertex-in-tile, uint16 peer slot});
 * peer_hw_a / peer_hw_b: HOST arrays of n_send_peers peer-mapped pointers to the peers' hw buffers of
 * sets A and B, peer_flag: to their flag slot for this rank; flags: this rank's uint32 flag array;
 * wait_ranks_dev: device int32[n_wait] source ranks.  Sweep i waits for flag value sweep_base + 1 + i and
 * raises sweep_base + 2 + i (the halo of the initial state is published with value sweep_base + 1, e.g.
 * by nxb_halo_put_f32).  ticket: device uint32, zero. */
int nxb_erode3_run_comm_f32(const void *plan_mem, const int32_t *adj, const float *dist, const float *dist3,
                            float *hw_a, float *s_a, float *hw_b, float *s_b,
                            int64_t n_own, float rain, int64_t n_sweeps,
                            const void *send_entries, int n_send_peers,
                            void *const *peer_hw_a, void *const *peer_hw_b, void *const *peer_flag,
                            const void *flags, const int32_t *wait_ranks_dev, int n_wait,
                            uint32_t sweep_base, void *ticket, void *stream);
/* Reference-exact mode of the sweep: float64 positions / state, no FMA, the reference's operation and
 * neighbour order -> bit-identical to erosion_iteration3 (erosion.py:197-279) preceded by
 * `water += rain` (erosion.py:182-183).  nodes: double[.][3]; ping-pong buffers of n doubles. */
int nxb_erode3_step_f64(const double *nodes, const int32_t *adj,
                        const double *h_in, const double *w_in, const double *s_in,
                        double *h_out, double *w_out, double *s_out, int64_t n, double rain, void *stream);
/* erosion.py:76-99 erosion_iteration1 */
int nxb_erode1_step_f32(const int32_t *adj, const float *h_in, float *h_out,
                        int64_t v_begin, int64_t v_end, void *stream);
/* ---- multi-GPU halo exchange over NVLink peer memory (csrc/nxb_halo.cu) ------------------------
 * Store this rank's boundary {height, water} pairs straight into every peer's halo slots (peer_hw[p] is
 * the peer's hw buffer mapped into this process, by torch.distributed._symmetric_memory or nxb_peer_*)
 * and raise flag_value in the peer's flag slot for this rank.  send_idx: concatenated LOCAL indices,
 * peer after peer (src_begin / count per peer); dst_off[p] = first element of peer p's buffer this rank
 * fills.  ticket: device uint32, zero.  One fused kernel; publishes the initial state of a run. */
int nxb_halo_put_f32(const float *hw, const int32_t *send_idx, int npeers,
                     void *const *peer_hw, void *const *peer_flag,
                     const int64_t *dst_off, const int64_t *src_begin, const int64_t *count,
                     uint32_t flag_value, void *ticket, void *stream);
/* Stream-ordered wait until flags[src_ranks[i]] >= target for all i (flags: this rank's uint32 array,
 * written by the peers' nxb_halo_put_f32; src_ranks: device int32[npeers]). */
int nxb_halo_wait(const void *flags, const int32_t *src_ranks, int npeers, uint32_t target, void *stream);
/* the same wait as stream memory operations (cuStreamWaitValue32, one per flag): no kernel, no SM.
 * src_ranks_host: HOST int32[npeers]. */
int nxb_halo_wait_stream(const void *flags, const int32_t *src_ranks_host, int npeers, uint32_t target, void *stream);

/* Peer-mapped buffers through CUDA IPC (the alternative to torch symmetric memory; also works for two
 * processes on ONE device): the owner allocates zeroed device memory and exports a 64-byte handle
 * (HOST buffer), peers open it and receive a device pointer valid in their process.  Synchronous. */
int nxb_peer_alloc(int64_t bytes, void **ptr_out, void *handle64_host);
int nxb_peer_open(const void *handle64_host, void **ptr_out);
int nxb_peer_close(void *ptr);
int nxb_peer_free(void *ptr);

/* gather / scatter of halo values for the multi-GPU exchange: dst[i] = src[idx[i]] and
 * dst[idx[i]] = src[i] */
int nxb_gather_f32(const float *src, const int32_t *idx, int64_t n, float *dst, void *stream);
int nxb_scatter_f32(const float *src, const int32_t *idx, int64_t n, float *dst, void *stream);

/* float <-> double conversion of result arrays (reference API is float64) */
int nxb_f32_to_f64(const float *src, int64_t n, double *dst, void *stream);
int nxb_f64_to_f32(const double *src, int64_t n, float *dst, void *stream);

#ifdef __cplusplus
}
#endif
#endif
