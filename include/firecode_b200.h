/* firecode_b200 -- C-ABI of the B200-native FIRECODE embedding screen.
 *
 * Every entry point is plain C: pointers + sizes, no torch / C++ types.  Functions return 0 on
 * success and a non-zero code on failure; fc_last_error() then holds a thread-local message.
 * "_dev" entry points take DEVICE pointers and a cudaStream_t passed as void* (0 = default
 * stream) and never synchronise; the un-suffixed ones take HOST pointers, do their own
 * host<->device copies on an internal stream and return when the results are in host memory.
 *
 * Reference citations are file:line under /root/reference/firecode (FIRECODE v2.0.4).
 */
#ifndef FIRECODE_B200_H
#define FIRECODE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FC_OK 0
#define FC_ERR_INVALID 1
#define FC_ERR_CUDA 2
#define FC_ERR_NOMEM 3

/* status byte written per pose by the clash screen */
#define FC_STATUS_PASS 1u      /* bit0: pose passes the compenetration check                       */
#define FC_STATUS_RECHECKED 2u /* bit1: decided by the FP64 recheck (FP32 value was inside the band) */
#define FC_STATUS_NEAR 4u      /* bit2: FP64 min distance within FC_NEAR_EPS of the threshold        */
#define FC_NEAR_EPS 1e-6

const char* fc_last_error(void);
int fc_version(void);
/* number of visible CUDA devices; <= 0 means the library cannot run (there is no CPU fallback) */
int fc_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * Clash screen (compenetration check).
 * Replaces utils.py:507-575 `compenetration_check(coords, ids=..., thresh, max_clashes)` as it is
 * called once per candidate pose at embeds.py:139-141, 559-563, 718-722, and torsion_module.py:
 * 894-918 `torsion_comp_check`.  A pose is (conformer a of ensemble A, conformer b of ensemble B,
 * rigid transform): fragment A stays in its own frame, fragment B is placed at R @ b + t, exactly
 * the `get_embed` expression embeds.py:815-817 with mol1 at identity (only the relative transform
 * matters for intermolecular distances).  pass <=> #{(i,j): |a_i - (R b_j + t)| < thresh} <=
 * max_clashes  (strict `<`, utils.py:551; `strict == 0` selects the `<=` of the trimolecular
 * branch utils.py:563-571).
 * ---------------------------------------------------------------------------------------------- */

/* Largest number of poses one tile may hold for fragments of n_b atoms (explicit tile lists must
 * not exceed it) and the FP32-kernel geometry chosen for it. */
int fc_clash_tile_poses(int n_b);

/* Device-pointer entry.
 *  a_coords  (n_conf_a, n_a, 3) f64   ensemble A
 *  b_coords  (n_conf_b, n_b, 3) f64   ensemble B
 *  xf        (n_poses, 12)      f64   per pose: R row-major (9) then t (3)
 *  tiles     (n_tiles, 4)       i32   {conf_a, conf_b, first_pose, n_poses_in_tile<=tile_poses} or
 *                                      NULL: every pose uses conformer 0 of both ensembles
 *  status    (n_poses)          u8    FC_STATUS_* bits
 *  min_dist  (n_poses)          f32   optional (may be NULL): FP32 estimate of the min distance
 *  near_count (1) i32 device, near_idx (near_cap) i64, near_dist (near_cap) f64: optional list of
 *            poses whose FP64 min distance lies within FC_NEAR_EPS of thresh (NULL to skip)
 *  pose_index_base  added to the pose numbers written to near_idx (chunked callers)
 * Scratch memory is taken from the stream-ordered pool (cudaMallocAsync) on `stream`.
 */
int fc_clash_screen_dev(const double* a_coords, int n_conf_a, int n_a, const double* b_coords,
                        int n_conf_b, int n_b, const double* xf, int64_t n_poses,
                        const int32_t* tiles, int64_t n_tiles, double thresh, int max_clashes,
                        int strict, uint8_t* status, float* min_dist, int32_t* near_count,
                        int64_t* near_idx, double* near_dist, int64_t near_cap,
                        int64_t pose_index_base, void* stream);

/* Fragment A's tables (FP32 atom-pair layout, bounding radii and -- with want_cells -- the cell grid,
 * candidate records and occupancy bits) can be built once and reused by several screens that share
 * fragment A and the threshold (the trimolecular embed screens every molecule against many chunks).
 * All calls are stream-ordered; free the preparation on the same stream after the last screen. */
typedef struct fc_clash_prep fc_clash_prep;
int fc_clash_prepare_dev(const double* a_coords, int n_conf_a, int n_a, double thresh, int want_cells,
                         fc_clash_prep** out, void* stream);
int fc_clash_screen_prepared_dev(const fc_clash_prep* prep, const double* a_coords, const double* b_coords,
                                 int n_conf_b, int n_b, const double* xf, int64_t n_poses,
                                 const int32_t* tiles, int64_t n_tiles, int max_clashes, int strict,
                                 uint8_t* status, float* min_dist, int32_t* near_count, int64_t* near_idx,
                                 double* near_dist, int64_t near_cap, int64_t pose_index_base, void* stream);
void fc_clash_prep_free(fc_clash_prep* prep, void* stream);
/* Test hook (host pointers): out8 = {x, y, z of the centre of cell (0,0,0), cell edge h, cells per axis, candidate
 * radius, 1 / ((g-1) h), log2 g} of the cell grid the screen builds for this ensemble and threshold (zeros when
 * it builds none); tests place atoms on cell faces with it. */
int fc_clash_cell_meta(const double* a_coords, int n_conf_a, int n_a, double thresh, float* out8);

/* General device entry of the screen: poses in either format, status bytes and / or the survivor bitmask.
 *  pose_format FC_POSE_XF64: poses = (n_poses, 12) f64 as above.
 *  pose_format FC_POSE_Q7  : poses = (n_poses, 7) f32 {qx, qy, qz, qw, tx, ty, tz}: any non-zero quaternion and
 *            a translation, 28 bytes per pose instead of 96.  The pose it stands for is DEFINED as the FP64 expansion
 *            (every operation rounded on its own, in this order; x, y, z, w, t converted to f64 first)
 *                n = ((x*x + y*y) + z*z) + w*w;   s = 2 / n
 *                R = [[1 - s*(y*y + z*z), s*(x*y - z*w),     s*(x*z + y*w)    ],
 *                     [s*(x*y + z*w),     1 - s*(x*x + z*z), s*(y*z - x*w)    ],
 *                     [s*(x*z - y*w),     s*(y*z + x*w),     1 - s*(x*x + y*y)]],   t = (tx, ty, tz)
 *            which is what the FP64 recheck evaluates and what a CPU oracle must expand before it applies the
 *            reference arithmetic (oracle/port.py:pose7_to_xf).
 *  status    (n_poses) u8 or NULL;  bits (ceil(n_poses / 32)) u32 or NULL (at least one of the two): bit (i & 31) of
 *            word i >> 5 = pose i passes -- written by the screen itself (warp ballot), the buffer the ranks exchange.
 *  recheck_count (1) i32 device or NULL: incremented by the number of poses the FP64 recheck decided.
 * Everything else as fc_clash_screen_prepared_dev. */
#define FC_POSE_XF64 0
#define FC_POSE_Q7 1
int fc_clash_screen_ex_dev(const fc_clash_prep* prep, const double* a_coords, const double* b_coords, int n_conf_b,
                           int n_b, const void* poses, int pose_format, int64_t n_poses, const int32_t* tiles,
                           int64_t n_tiles, int max_clashes, int strict, uint8_t* status, uint32_t* bits,
                           float* min_dist, int32_t* near_count, int64_t* near_idx, double* near_dist, int64_t near_cap,
                           int64_t pose_index_base, int32_t* recheck_count, void* stream);

/* Host-pointer entry for compact poses (FC_POSE_Q7): 28 bytes per pose go to the device and one BIT per pose comes
 * back (bits_out, ceil(n_poses / 32) words; status_out optional), in chunks pipelined with the kernels.  Pass pinned
 * memory (fc_host_alloc) for pose7 to reach the PCIe rate.  counts as fc_clash_batch. */
int fc_clash_batch_pose7(const double* a_coords, int n_conf_a, int n_a, const double* b_coords, int n_conf_b, int n_b,
                         const float* pose7, int64_t n_poses, const int32_t* tiles, int64_t n_tiles, double thresh,
                         int max_clashes, int strict, uint32_t* bits_out, uint8_t* status_out, int64_t* counts,
                         int64_t* near_idx, double* near_dist, int64_t near_cap);

/* Page-locked host memory for the host-pointer entry points (cudaHostAlloc), placed on the NUMA node of the current
 * device when the kernel allows it (set_mempolicy(MPOL_PREFERRED) around the allocation; the node comes from
 * /sys/bus/pci/devices/<gpu>/numa_node).  node_out (may be NULL) receives that node or -1. */
void* fc_host_alloc(int64_t bytes, int32_t* node_out);
void fc_host_free(void* p);

/* Host-pointer entry: same arguments in host memory; copies are pipelined in chunks with the
 * kernels.  counts[0] = passing poses, counts[1] = poses decided by the FP64 recheck,
 * counts[2] = near-threshold poses (near_idx/near_dist hold the first near_cap of them). */
int fc_clash_batch(const double* a_coords, int n_conf_a, int n_a, const double* b_coords,
                   int n_conf_b, int n_b, const double* xf, int64_t n_poses, const int32_t* tiles,
                   int64_t n_tiles, double thresh, int max_clashes, int strict, uint8_t* status,
                   float* min_dist, int64_t* counts, int64_t* near_idx, double* near_dist,
                   int64_t near_cap);

/* FP32-kernel geometry chosen for fragments of n_b atoms: out4 = {atoms per thread, threads per
 * pose, poses per tile, threads per block}. */
int fc_clash_geometry(int n_b, int32_t* out4);

/* CUDA-event timing of the FP32 clash kernel alone (bench.py roofline leg): returns the summed
 * milliseconds and launch count recorded on this thread since the previous call, then switches
 * recording on (enable != 0) or off. */
int fc_clash_timing(int enable, double* ms_sum, int64_t* launches);

/* Survivor bitmask: bit (i & 31) of bits[i >> 5] = status[i] & FC_STATUS_PASS; bits holds
 * ceil(n / 32) words.  This is the buffer the ranks exchange with one NCCL all-gather. */
int fc_pack_mask_dev(const uint8_t* status, int64_t n, uint32_t* bits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Embed screens.  Results come back in an opaque fc_result (sizes are not known in advance).
 * ---------------------------------------------------------------------------------------------- */
#define FC_TIE_CLASH 1  /* min distance within FC_NEAR_EPS of the clash threshold            */
#define FC_TIE_TFD 2    /* torsion-difference sum within FC_NEAR_EPS of its threshold        */
#define FC_TIE_RMSD 3   /* RMSD within FC_NEAR_EPS of its threshold                          */
#define FC_TIE_MAXDEV 4 /* max deviation within FC_NEAR_EPS of its threshold                 */

/* one near-threshold decision: `a` is a pose / structure index, `b` the earlier pose / structure of
 * the pair (-1 for clash decisions), `decision` = 1 when the CUDA path evaluated value < threshold */
typedef struct fc_tie {
    int64_t a, b;
    double value;
    int32_t kind, decision;
} fc_tie;

typedef struct fc_result fc_result;
void fc_result_free(fc_result* r);
/* out10 = {poses screened, clash survivors, FP64 rechecks, kept, near-threshold decisions,
 *          atoms per pose, stage-1 survivors, quadruplets, constrained pairs per pose, groups} */
int fc_result_counts(const fc_result* r, int64_t* out10);
int fc_result_status(const fc_result* r, uint8_t* out);         /* (poses screened) FC_STATUS_* */
int fc_result_survivors(const fc_result* r, int64_t* out);      /* (survivors) pose indices      */
int fc_result_fingerprints(const fc_result* r, double* out);    /* (survivors, quadruplets) deg  */
int fc_result_kept_indices(const fc_result* r, int64_t* out);   /* (kept) pose indices, in order */
int fc_result_kept_coords(const fc_result* r, double* out);     /* (kept, atoms per pose, 3)     */
int fc_result_constrained(const fc_result* r, int32_t* out);    /* (kept, n_pairs, 2)            */
int64_t fc_result_ties(const fc_result* r, fc_tie* out, int64_t cap); /* returns total recorded  */
/* trimolecular embeds: per group the grid-search candidate (0..342) chosen by the restated
 * _adjust_directions (embeds.py:403-405) and the cost gap (degrees) to the runner-up */
int fc_result_groups(const fc_result* r, int32_t* choice, double* gap);

/* String embed: replaces embeds.py:51-158 `string_embed(embedder)`.
 * Pose index = enumeration order of the reference loops (conformer pairs in cartesian_product
 * order, utils.py:219-221, then orbital-centre pairs, then angles). */
typedef struct fc_string_problem {
    const double* coords1;  int32_t n_conf1, n_atoms1;  /* (n_conf1, n_atoms1, 3) Hypermolecule.coords */
    const double* coords2;  int32_t n_conf2, n_atoms2;
    const double* centers1; const double* vecs1; int32_t k1; /* (n_conf1, k1, 3) RAtom.center / orb_vecs */
    const double* centers2; const double* vecs2; int32_t k2;
    const double* angles;   int32_t n_angles;               /* embedder.systematic_angles, degrees      */
    const int64_t* quadruplets; int32_t n_quads;            /* (n_quads, 4) torsion_module.py:385-408   */
    double thresh;          /* options.clash_thresh */
    int32_t max_clashes;    /* 0 in the embeds (embeds.py:139-141) */
    int32_t rot_handedness; /* +1 / -1, prism_pruner rot_mat_from_pointer convention */
    double tfd_thresh;      /* 10 degrees, torsion_module.py:1056 */
} fc_string_problem;

int64_t fc_string_n_poses(const fc_string_problem* p);
/* whole screen on the current device */
int fc_string_screen(const fc_string_problem* p, fc_result** out);
/* sharded form: (1) clash screen + torsion fingerprints of poses [pose_lo, pose_hi) on this rank,
 * (2) after the ranks all-gathered survivors and fingerprints, the ordered keep-first sweep,
 * (3) coordinates of the kept poses. */
int fc_string_stage1(const fc_string_problem* p, int64_t pose_lo, int64_t pose_hi, fc_result** out);
int fc_tfd_keepfirst(const double* fingerprints, const int64_t* labels, int64_t n, int32_t n_quads,
                     double thresh, uint8_t* keep_out, fc_tie* ties_out, int64_t tie_cap,
                     int64_t* n_ties_out);
int fc_string_materialize(const fc_string_problem* p, const int64_t* kept, int64_t n_kept, double* out);

/* Cyclical embed, bimolecular path: replaces embeds.py:588-750 `_fast_bimol_rigid_cyclical_embed`
 * (what `cyclical_embed` runs for two molecules, embeds.py:184-185, "cyclical" and "chelotropic").
 * A group is one (conformer pair, pivot pair, polygon orientation) combination that passed the
 * reference's norm and pairing filters; the host enumerates groups in the reference's loop order
 * (O(conformers x pivots) work) and the GPU expands every group into n_angles poses.
 * Pose index = group * n_angles + angle index. */
typedef struct fc_cyclical_problem {
    int32_t n_mols;                       /* 2 */
    const double* coords[3];  int32_t n_conf[3], n_atoms[3];
    const int64_t* reactive[3]; int32_t n_reactive[3];   /* Hypermolecule.reactive_indices (1 or 2) */
    int64_t n_groups;
    const int32_t* group_conf;   /* (n_groups, n_mols) conformer of every molecule                 */
    const double* group_pivot;   /* (n_groups, n_mols, 3) Pivot.pivot of the active pivot           */
    const double* group_mean;    /* (n_groups, n_mols, 3) Pivot.meanpoint                           */
    const double* group_vecs;    /* (n_groups, n_mols, 2, 3) polygonize start/end for orientation v */
    const double* group_dirs;    /* (n_groups, n_mols, 3) alignment directions                      */
    const int32_t* group_ids;    /* (n_groups, n_pairs, 2) atom couples to constrain (may be NULL)  */
    int32_t n_pairs;
    const double* angles; int32_t n_angles;   /* (n_angles, n_mols) embedder.systematic_angles      */
    double thresh; int32_t max_clashes; int32_t rot_handedness;
    double rmsd_thresh;          /* 1.0 in the embeds (embeds.py:724) */
} fc_cyclical_problem;

int fc_cyclical_screen(const fc_cyclical_problem* p, fc_result** out);

/* Host only (no CUDA call): the group table fc_cyclical_screen consumes, enumerated in the reference's loop order
 * (embeds.py:596-641): conformer pairs (first index fastest, utils.py:219-221), pivot pairs (first index fastest), two
 * orientations; pivot pairs whose norms differ by more than max_norm_delta are dropped (embeds.py:624), and so are the
 * arrangements that miss a user pairing (embeds.py:638-641; a pairing found in `internal` counts as satisfied).
 *  n_conf[2]; per molecule m the pivot tables are CSR over its conformers: rows off_m[c] .. off_m[c+1] of
 *  vec_m (rows, 3) Pivot.pivot, mean_m (rows, 3) Pivot.meanpoint, ids_m (rows, 2) start / end cumnum.
 *  Writes at most cap groups into conf (G,2) pivot (G,2,3) mean (G,2,3) vecs (G,2,2,3) dirs (G,2,3) ids (G,2,2) and
 *  returns their number in *n_groups (larger than cap: call again). */
int fc_cyclical_groups(const int32_t* n_conf, const int64_t* off0, const double* vec0, const double* mean0,
                       const int64_t* ids0, const int64_t* off1, const double* vec1, const double* mean1,
                       const int64_t* ids1, double max_norm_delta, const int64_t* pairings, int32_t n_pairings,
                       const int64_t* internal, int32_t n_internal, int64_t cap, int64_t* n_groups, int32_t* conf_out,
                       double* pivot_out, double* mean_out, double* vecs_out, double* dirs_out, int32_t* ids_out);

/* Cyclical embed, trimolecular body: replaces embeds.py:409-585 (`cyclical_embed` for three
 * molecules) including `_get_directions` (embeds.py:188-254) and the stateful 343-point grid search
 * `_adjust_directions` (embeds.py:256-407).  The library enumerates (conformer triple, pivot
 * triple) super-groups in the reference's loop order (cartesian_product order, utils.py:219-221:
 * third index fastest, then the first, the second outermost), drops impossible triangles
 * (embeds.py:447-462), applies the pairing filter per polygon orientation (embeds.py:473-476) and
 * expands every surviving (super-group, orientation) group into n_angles poses.
 * Pose index = group * n_angles + angle index, groups numbered in loop order.
 * Clash test = the three-block `<=` branch utils.py:553-575 with max_clashes = 0; because block
 * (i, j) only depends on the angles of molecules i and j, each block is screened once per distinct
 * (angle_i, angle_j) pair of the angle table (36 instead of 216 for the default 6x6x6 grid).
 * Pivot tables are CSR over conformers: rows pivot_offsets[m][c] .. pivot_offsets[m][c+1]. */
typedef struct fc_cyclical3_problem {
    const double* coords[3];  int32_t n_conf[3], n_atoms[3];
    const int64_t* reactive[3]; int32_t n_reactive[3];   /* Hypermolecule.reactive_indices (1 or 2)  */
    const int64_t* pivot_offsets[3];  /* (n_conf + 1)                                                */
    const double* pivot_vec[3];       /* (rows, 3) Pivot.pivot                                       */
    const double* pivot_mean[3];      /* (rows, 3) Pivot.meanpoint                                   */
    const int64_t* pivot_ids[3];      /* (rows, 2) start_atom.cumnum, end_atom.cumnum                */
    const int64_t* ratoms0[3]; int32_t n_ratoms0[3];  /* (k, 2) {index, cumnum} of conformer 0's
                                                         reactive atoms, dict order (embeds.py:330)  */
    const double* angles; int32_t n_angles;           /* (n_angles, 3) embedder.systematic_angles    */
    const int64_t* pairings; int32_t n_pairings;      /* (n, 2) embedder.pairings_table.values()     */
    const int64_t* internal; int32_t n_internal;      /* (n, 2) internal constraints a pairing may
                                                         match (empty for the ndarray quirk N10)     */
    double thresh; int32_t rot_handedness; double rmsd_thresh;
    int64_t conf_tuple_lo, conf_tuple_hi; /* slice of the conformer-triple enumeration (hi <= 0: all) */
    int32_t flags;                        /* FC_CYC3_* */
} fc_cyclical3_problem;
#define FC_CYC3_NO_STATUS 1 /* do not return the per-pose status bytes  */
#define FC_CYC3_NO_COORDS 2 /* do not materialise the kept poses        */

int fc_cyclical3_screen(const fc_cyclical3_problem* p, fc_result** out);

/* Ensemble similarity pruning: replaces prism_pruner.pruner.prune_by_rmsd (mode 0) and
 * prune_by_moment_of_inertia (mode 1) as called at embedder.py:1452,1472; ensemble.py:211,230;
 * operators.py:613-624; atropisomer_module.py:504.  Multi-pass chunked driver (k = 500000 ... 1,
 * a pass runs when k == 1 or min_per_chunk * k < active structures; chunks of n / k structures).
 *  structures (n, n_atoms, 3) f64 host;  sel (n_sel) atoms entering the RMSD (heavy atoms);
 *  masses (n_atoms) for mode 1;  energies (n) or NULL: pairs with |dE| >= max_dE are not compared;
 *  keep_first / snapshot: the unpinned prism_pruner conventions; `snapshot` is a flag word: bit 0
 *  (FC_PRUNE_SNAPSHOT) a pass reads the mask as it was when the pass started, bit 1 (FC_PRUNE_CHUNK_ACTIVE) the k
 *  chunks of a pass hold n_active / k consecutive ACTIVE structures instead of n / k consecutive structures;
 *  mask_out (n) 1 = kept;  stats_out[4] = {passes, pairs whose covariance was accumulated, pairs
 *  eigen-solved, pairs skipped because both structures survived an earlier pass in the same chunk
 *  (such survivors are mutually dissimilar by construction)}.  Similar pairs are collected in a
 *  compact list; the order-dependent keep rule of a pass is resolved on the host in O(list).
 *  Mode 0 screens the pairs on the tensor cores (TF32 Gram matrix of the centred coordinates, rigorous
 *  rounding band) when n_sel <= 88 and on the FP32 CUDA cores otherwise; every pair a screen cannot rule
 *  out is decided in FP64, so mask_out does not depend on the screen. */
#define FC_PRUNE_SNAPSHOT 1
#define FC_PRUNE_CHUNK_ACTIVE 2
int fc_prune(const double* structures, int64_t n, int32_t n_atoms, int32_t mode, const int32_t* sel,
             int32_t n_sel, const double* masses, double max_rmsd, double max_dev, double moi_dev,
             const double* energies, double max_dE, int32_t keep_first, int32_t snapshot,
             int32_t min_per_chunk, uint8_t* mask_out, int64_t* stats_out, fc_tie* ties_out,
             int64_t tie_cap, int64_t* n_ties_out);

/* Multi-GPU form: the pair tiles of every pass are dealt round-robin to `world` ranks (structures
 * replicated); the similar pairs found by each rank are all-gathered through `gather` -- the host
 * language implements it with NCCL / gloo -- and every rank resolves the pass on the union, so all
 * ranks return the same mask.  gather(send, send_bytes, &recv, &recv_bytes, ctx) must return 0 and
 * hand back the rank-order concatenation of all ranks' buffers (owned by the callee until the next
 * call).  stats_out counts this rank's share. */
typedef int (*fc_allgather_fn)(const void* send, int64_t send_bytes, const void** recv, int64_t* recv_bytes, void* ctx);
int fc_prune_sharded(const double* structures, int64_t n, int32_t n_atoms, int32_t mode, const int32_t* sel,
                     int32_t n_sel, const double* masses, double max_rmsd, double max_dev, double moi_dev,
                     const double* energies, double max_dE, int32_t keep_first, int32_t snapshot,
                     int32_t min_per_chunk, uint8_t* mask_out, int64_t* stats_out, fc_tie* ties_out,
                     int64_t tie_cap, int64_t* n_ties_out, int32_t rank, int32_t world,
                     fc_allgather_fn gather, void* gather_ctx);

/* The same with the exchange steps on the GPUs (NCCL over NVLink behind the host language).  Every rank uploads only
 * its 1 / world of the structures (rows [rank * ceil(n / world), ...)); the pieces, the per-pass counters and the
 * similar-pair lists are exchanged device to device through
 *   gather_dev(send, recv, bytes, stream, ctx): every rank contributes `bytes` bytes at the DEVICE address `send`;
 *   on return the work is ordered on `stream` (a cudaStream_t) and leaves the rank-order concatenation (world * bytes)
 *   at the DEVICE address `recv`.  Returns 0 on success.
 * The union of a pass is sorted on the device and every rank resolves the same list: all ranks return the same mask,
 * equal to the single-GPU one.  BASELINE config C4 "sharded over 8 GPUs". */
typedef int (*fc_allgather_dev_fn)(const void* send_dev, void* recv_dev, int64_t bytes, void* stream, void* ctx);
int fc_prune_sharded_dev(const double* structures, int64_t n, int32_t n_atoms, int32_t mode, const int32_t* sel,
                         int32_t n_sel, const double* masses, double max_rmsd, double max_dev, double moi_dev,
                         const double* energies, double max_dE, int32_t keep_first, int32_t snapshot,
                         int32_t min_per_chunk, uint8_t* mask_out, int64_t* stats_out, fc_tie* ties_out,
                         int64_t tie_cap, int64_t* n_ties_out, int32_t rank, int32_t world,
                         fc_allgather_dev_fn gather_dev, void* gather_ctx);

/* Batched bond-graph checks: the post-filters `scramble_check` and `molecule_check` (utils.py:341-400; callers
 * embedder.py:2181-2195, 2430-2444, optimization_methods.py:139-147, operators.py:502, interfaces/goat.py:309).
 * Atoms i != j of a structure are bonded iff |x_i - x_j| < factor * (radii[i] + radii[j]) -- prism_pruner's
 * graphize / d_min_bond criterion (factor 1.2, covalent radii), FP64, strict <.
 *  fc_bond_graph_batch: adj_out (n_struct, n_atoms, W = ceil(n_atoms / 32)) words, bit (j & 31) of word j >> 5 of row i
 *    set iff i - j bonded (symmetric, no self loops).
 *  fc_bond_delta_batch: delta_out[s] = number of bonds present in exactly one of {structure s, expected}, not counting
 *    bonds that touch an atom with excluded[i] != 0 (excluded may be null).  `expected` has the layout of adj_out for
 *    ONE structure (expected_per_structure = 0: scramble_check's union of the fragment graphs) or for every structure
 *    (1: molecule_check against each structure's own un-optimised graph).
 *  near_out[s] (may be null) = pairs of structure s whose distance lies within 1e-6 of their limit. */
int fc_bond_graph_batch(const double* coords, int64_t n_struct, int32_t n_atoms, const double* radii, double factor,
                        uint32_t* adj_out, int32_t* near_out);
int fc_bond_delta_batch(const double* coords, int64_t n_struct, int32_t n_atoms, const double* radii, double factor,
                        const uint32_t* expected, int32_t expected_per_structure, const uint8_t* excluded,
                        int32_t* delta_out, int32_t* near_out);

/* Torsion-fingerprint (TFD) ensemble pruning, the O(n^2) part of torsion_module.py:957-1043
 * `prune_conformers_tfd` (embedder.py:1430-1437, torsion_module.py:875).
 *  fc_tfd_fingerprints: `_get_tf_mat` (torsion_module.py:1046-1053): tf_out (n, n_quads) degrees.
 *  fc_tfd_first_match : for every structure i of every chunk {start, len} of a pass, first_out[i] = the
 *  first later structure j of the same chunk with sum_q wrap180(|tf_i - tf_j|) < thresh
 *  (torsion_module.py:1056-1067), -1 if none -- what the reference's pair loops record before they break.
 *  The host (Python, same networkx calls as the reference) turns the matches into clusters. */
int fc_tfd_fingerprints(const double* structures, int64_t n, int32_t n_atoms, const int64_t* quadruplets,
                        int32_t n_quads, double* tf_out);
int fc_tfd_first_match(const double* tf, int64_t n, int32_t n_quads, const int64_t* chunk_start,
                       const int64_t* chunk_len, int64_t n_chunks, double thresh, int64_t* first_out,
                       fc_tie* ties_out, int64_t tie_cap, int64_t* n_ties_out);

/* Batched torsion rotation with clash filtering: replaces the primitive pair
 * prism_pruner.utils.rotate_dihedral + torsion_module.py:894-918 `torsion_comp_check` (used at
 * torsion_module.py:523-552, 813-856) over every (conformer, torsion, angle) item.
 *  coords (n_conf, n_atoms, 3); torsions (n_tors, 4) i1..i4; masks (n_tors, n_atoms) 1 = atom moves
 *  (torsion_module.py:354-382); angles (n_angles) degrees.  Item order: conformer, torsion, angle.
 *  out_coords (items, n_atoms, 3) optional; status_out (items) FC_STATUS_PASS | FC_STATUS_NEAR;
 *  min_dist_out (items) optional: smallest moved-static distance. */
int fc_torsion_scan(const double* coords, int32_t n_conf, int32_t n_atoms, const int32_t* torsions,
                    int32_t n_tors, const uint8_t* masks, const double* angles, int32_t n_angles,
                    double thresh, int32_t max_clashes, int32_t rot_handedness, int32_t axis_sign,
                    double* out_coords, uint8_t* status_out, double* min_dist_out);

/* prism_pruner.rmsd.rmsd_and_max(ref, structure, center) for n structures against one reference: the
 * arithmetic of utils.py:494-504 `rmsd_similarity` (center = 0, all atoms) and embedder.py:1784-1786
 * (center = 1).  ref (n_atoms, 3), structures (n, n_atoms, 3) host; outputs (n). */
int fc_rmsd_and_max_batch(const double* ref, const double* structures, int64_t n, int32_t n_atoms,
                          int32_t center, double* rmsd_out, double* maxdev_out);

/* Symmetry-corrected RMSD of structure pairs, the per-pair arithmetic of prism_pruner.pruner.prune_by_rmsd_rot_corr
 * (call sites embedder.py:1485-1496, ensemble.py:253, operators.py:626; algorithm [UNVERIFIED-RECALL], stated in
 * csrc/fc_rotcorr.cu and oracle/prism_pruner/pruner.py): for pair p = {reference r, structure c}, a copy of c has every
 * symmetric torsion, in order, set to the symmetry angle whose rotation of atom i4 alone best matches r on the four
 * torsion atoms (first minimum), the torsion's rotating group (masks) following; then
 * rmsd_and_max(r[sel], copy[sel], center=True).
 *  structures (n, n_atoms, 3) host; torsions (n_tors, 4); masks (n_tors, n_atoms); angles flat, torsion t owns
 *  angles[angle_offsets[t] .. angle_offsets[t+1]) degrees; pairs (n_pairs, 2); outputs (n_pairs);
 *  choice_out / gap_out (n_pairs, n_tors) optional: winning angle index per torsion and the RMSD gap to the runner-up. */
int fc_rmsd_rot_corr_pairs(const double* structures, int64_t n, int32_t n_atoms, const int32_t* sel, int32_t n_sel,
                           const int32_t* torsions, int32_t n_tors, const uint8_t* masks, const double* angles,
                           const int32_t* angle_offsets, const int32_t* pairs, int64_t n_pairs, int32_t rot_handedness,
                           int32_t axis_sign, double* rmsd_out, double* maxdev_out, int32_t* choice_out, double* gap_out);

/* Non-fragment branch of utils.py:523-542 `compenetration_check(coords, graph)` and algebra.py:52-54
 * `count_clashes` for n structures: close_pairs_out[s] = ordered atom pairs with 0 < d < 0.5 A,
 * nonbonded_out[s] = ordered pairs i != j with d < thresh and bonded[i * n_atoms + j] == 0
 * (bonded may be NULL). */
int fc_self_clash_batch(const double* coords, int64_t n, int32_t n_atoms, const uint8_t* bonded, double thresh,
                        int64_t* close_pairs_out, int64_t* nonbonded_out);

/* Conformational-search inner loop, torsion_module.py:512-552 (random_csearch) = 813-856
 * (clustered_csearch): for every (starting structure, angle set) the torsions of the set are applied in
 * order -- rotate_dihedral, torsion_comp_check (thresh, max_clashes = 0), the 5-degree back-off loop
 * (at most angle // 5 steps) -- to the running coordinates.  Item order: start-major, then angle set.
 *  starts (n_starts, n_atoms, 3); torsions (n_tors, 4); masks (n_tors, n_atoms); angle_sets (n_sets, n_tors)
 *  integer degrees; out_coords (items, n_atoms, 3); rotated_out (items) number of bonds that rotated;
 *  near_out (items) 1 if a distance met on the way was within FC_NEAR_EPS of thresh. */
int fc_csearch_apply(const double* starts, int32_t n_starts, int32_t n_atoms, const int32_t* torsions,
                     int32_t n_tors, const uint8_t* masks, const int32_t* angle_sets, int64_t n_sets,
                     double thresh, int32_t rot_handedness, int32_t axis_sign, double* out_coords,
                     int32_t* rotated_out, uint8_t* near_out);

/* compenetration_check(structure, ids, thresh, max_clashes) (utils.py:544-575) for n complete structures --
 * the loop of RunEmbedding.compenetration_refining (embedder.py:1954-1975).  ids: atoms per fragment (2 or 3).
 * count_out[s] = clashing pairs (two fragments: d < thresh over (m2, m1); three: d <= thresh summed over
 * (m2,m1), (m3,m2), (m1,m3)); the structure passes iff count <= max_clashes.  closest_out[s] = min |d - thresh|. */
int fc_structure_clash_batch(const double* coords, int64_t n, int32_t n_atoms, const int32_t* ids, int32_t n_ids,
                             double thresh, int64_t* count_out, double* closest_out);

/* fitness_check over a batch (optimization_methods.py:163-180, the loop of RunEmbedding.fitness_refining,
 * embedder.py:1997-2039): error_out[s] = sum over the constraints of structure s of (|x_a - x_b| - target);
 * pairs (n, n_constraints, 2), targets (n, n_constraints) with NaN where the reference has None.  The structure
 * is kept iff error < threshold (decided by the caller, which also lists errors within 1e-6 of the threshold). */
int fc_fitness_batch(const double* coords, int64_t n, int32_t n_atoms, const int32_t* pairs, const double* targets,
                     int32_t n_constraints, double* error_out);

/* Host-only test hook (no CUDA call): the Kabsch rotation the kernels compute from a 3x3 cross-covariance h
 * (row-major): r_out = U D V^T of the SVD h = U S V^T with D = diag(1, 1, det(U V^T)) (prism_pruner
 * rmsd.get_alignment_matrix / algebra.py:42-49), sig_out (may be NULL) = singular values, the third signed by det(h).
 * Defined for rank-deficient h as well (collinear or single atoms): any optimal rotation, never NaN. */
int fc_kabsch_host(const double* h, double* r_out, double* sig_out);

/* Host-only (no CUDA call): the padded position list and the work items (row0, col_tile0, n_col_tiles, pend) the
 * tensor-core screen of fc_prune would use for one pass with k chunks over the structures with mask != 0, after a
 * pass with prev_k chunks (0: first pass).  counts_out = {pairs in this rank's items, pairs known dissimilar}.
 * Test hook for the planner (tests/test_host_logic.py); sizes come back in n_spos / n_work also when the buffers
 * are too small. */
int fc_prune_plan(const uint8_t* mask, int64_t n, int64_t k, int64_t prev_k, int32_t world, int32_t rank, int32_t n_sms,
                  int32_t chunk_over_active, int32_t* spos_out, int64_t spos_cap, int64_t* n_spos, int32_t* work_out,
                  int64_t work_cap, int64_t* n_work, int64_t* counts_out);

/* Host-only test hook: segment number of every padded position of the pass fc_prune_plan describes.  A segment is a run
 * of positions whose structures shared a chunk of the previous pass (the whole chunk when prev_k = 0); no pair inside a
 * segment is evaluated (first pass: only by position order), so the driver may order its positions freely -- it sorts
 * them by sigma_1 on the device for the tile culling.  Numbers do not decrease along the positions; padding carries the
 * number of the segment before it. */
int fc_prune_plan_segments(const uint8_t* mask, int64_t n, int64_t k, int64_t prev_k, int32_t chunk_over_active,
                           int32_t* seg_out, int64_t seg_cap, int64_t* n_seg);

/* Timing of the last fc_prune / fc_prune_sharded call on this thread (bench.py's roofline of the tensor-core
 * screen): out6 = {wall ms of the call, CUDA-event ms summed over the screen kernel launches, launches,
 * pair slots the screen evaluated (2048 per 128 x 16 tile), candidates it passed on, atoms per structure}. */
int fc_prune_timing(double* out6);
/* Tile statistics of the tensor-core screen of the same call: out2 = {128 x 16 tiles planned, tiles multiplied}; the
 * difference was skipped by the shape bound  E >= sum_k (sigma_k(P) - sigma_k(Q))^2  (singular values of the centred
 * coordinates) after the positions of every known-dissimilar segment had been sorted by sigma_1 (FC_PRUNE_CULL=0
 * disables both). */
int fc_prune_tiles(double* out2);

/* dst[j] = the j-th row of src with mask != 0 (row_bytes each), copied by several host threads: the
 * `structures[mask]` every pruning entry point returns (consumer: apply_mask, embedder.py:1400-1408).
 * n_dst must equal the number of selected rows. */
int fc_take_rows(const void* src, int64_t row_bytes, const uint8_t* mask, int64_t n, void* dst, int64_t n_dst);

/* xyz text of n structures, byte-identical to the reference's write_xyz (utils.py:105-116): per structure
 * "<n_atoms>\n<title>\n" and one line "%s     % .6f % .6f % .6f\n" per atom.  symbols = n_atoms fixed-width
 * (sym_stride bytes, NUL-padded) element symbols; titles = n NUL-terminated strings back to back (null: "temp").
 * *out_len receives the size of the text; FC_ERR_INVALID when out_cap is smaller (call again with that size). */
int fc_xyz_format(const char* symbols, int32_t sym_stride, const double* coords, int64_t n, int32_t n_atoms,
                  const char* titles, char* out, int64_t out_cap, int64_t* out_len);

/* FP32 FMA-pipe peak probe used by bench.py for the roofline denominator: runs a dependent-free
 * FFMA2 loop on every SM and returns achieved TFLOP/s (2 flop per FMA lane). */
int fc_probe_fp32_peak(double* tflops_out, double* ms_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FIRECODE_B200_H */
