/* apt_b200.h -- C ABI of libapt_b200.so, the B200 (sm_100a) species-assignment engine.
 *
 * APOGeT (the reference) has no FFI layer: its "operator API" is the C++ template
 * phylogeny::PhylogeneticTree<GENOME,UDATA> (src/core/tree/phylogenetictree.hpp:44-45) plus the
 * genome operators distance()/compatibility()/genealogy() (:441-442, :467-468, :276).  This header
 * is the boundary a reference-side binding would call for genomes that have a fixed-length device
 * encoding; include/apt/phylogenetictree.hpp is that binding (same class name, members and
 * semantics as the reference's template) and INTEGRATION.md shows how a host program switches.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero
 * apt_status otherwise, with the message available from apt_last_error() (thread-local);
 * no exception crosses the boundary.  All calls are synchronous at return unless named *_async.
 * Pointers are HOST pointers unless the parameter name starts with d_ (device memory of the
 * tree's CUDA device).  `stream` parameters are CUstream / cudaStream_t handles passed as void*.
 * One host thread drives one tree (the reference is single-threaded and non re-entrant).
 *
 * Genome encoding: `words` 32-bit words of payload per genome and three floats
 * (mu, sigma_in, sigma_out) = the BOC compatibility genes of src/core/crossover.h:43-49.
 *   APT_BITS: payload is a bit-string, distance = Hamming distance (integer, exact);
 *   APT_F32 : payload is `words` IEEE floats, distance = L2 norm of the difference.
 * compatibility(d) = exp(-(d-mu)^2 / (2 sigma^2)), sigma = sigma_in if d < mu else sigma_out
 * (crossover.h:76-92).
 */
#ifndef APT_B200_H
#define APT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APT_INVALID 0xFFFFFFFFu      /* SID::INVALID / GID::INVALID, treetypes.h:24-26, 74-76 */

enum apt_kind { APT_BITS = 0, APT_F32 = 1 };

enum apt_status {
  APT_OK = 0,
  APT_ERR_ARGUMENT = 1,     /* bad parameter (also: unknown species id, as utils::Thrower at :169-174) */
  APT_ERR_CUDA = 2,         /* CUDA runtime failure, or no sm_100 device */
  APT_ERR_STATE = 3,        /* reference invariant violated (the reference asserts or crashes there) */
  APT_ERR_NOMEM = 4
};

/* Event stream = the reference's Callbacks_t hooks (callbacks.hpp:16-45), in firing order. */
enum apt_event_type {
  APT_EV_STEPPED = 0,            /* a = step, b = number of alive species          (:261) */
  APT_EV_NEW_SPECIES = 1,        /* a = parent sid or APT_INVALID, b = new sid       (:422-423) */
  APT_EV_ENTERS_ENVELOPPE = 2,   /* a = sid, b = gid                                 (:674, :705) */
  APT_EV_LEAVES_ENVELOPPE = 3,   /* a = sid, b = gid                                 (:704) */
  APT_EV_MAJOR_CONTRIBUTOR = 4   /* a = sid, b = old parent sid, c = new parent sid  (:765) */
};
typedef struct apt_event { uint32_t type, a, b, c; } apt_event;

/* config::PTree (src/core/ptreeconfig.h:17-69, defaults ptreeconfig.cpp:7-19) + device choice. */
typedef struct apt_params {
  uint32_t kind;               /* apt_kind */
  uint32_t words;              /* payload words per genome (1..1024) */
  uint32_t rset_size;          /* rsetSize, enveloppe size K (1..32), default 5 */
  int32_t  full_continuous;    /* DEBUG_FULL_CONTINUOUS: 1 = avg-compat score (:456-478), 0 = vote count (:430-452) */
  int32_t  env_crit;           /* DEBUG_ENV_CRIT: 0 maxAverage, 1 maxMinDist (default), 2 maxAvgMinStdDev, 3 maxWeightedDist2Avg */
  double   compat_thr;         /* compatibilityThreshold    .1  */
  double   sim_thr;            /* similarityThreshold       .5  */
  double   avg_thr;            /* avgCompatibilityThreshold .25 */
  uint32_t trim_period;        /* stillbornTrimmingPeriod   100 */
  float    trim_thr;           /* stillbornTrimmingThreshold .25 */
  float    trim_delay;         /* stillbornTrimmingDelay    4   */
  uint32_t trim_min_delay;     /* stillbornTrimmingMinDelay 200 */
  int32_t  device;             /* CUDA device ordinal */
  uint32_t flags;              /* apt_flags */
} apt_params;

enum apt_flags {
  APT_FLAG_NO_EVENTS = 1u,       /* do not record the event stream */
  APT_FLAG_NO_CHILD_KERNEL = 2u, /* debug: sub-species candidates are scored by the host repair path only */
  APT_FLAG_FORCE_GENERIC = 4u,   /* debug: never use the compatibility look-up table */
  APT_FLAG_NO_STREAM_KERNEL = 8u,/* debug: single-parent births are resolved by the host walk over speculative scores */
  APT_FLAG_NO_FAMILY_KERNEL = 16u /* debug: births rejected at top level always walk the sub-species on the host */
};

/* Fills *p with the reference's compiled-in defaults (ptreeconfig.cpp:7-19). */
void apt_params_default (apt_params *p);

const char* apt_last_error (void);

/* 1 when a CUDA device with compute capability 10.x is visible, else 0 (never fails). */
int apt_device_available (void);

/* ------------------------------------------------------------------------------------------------
 * Tree engine: PhylogeneticTree<GENOME,UDATA> for device-encoded genomes.
 * ---------------------------------------------------------------------------------------------- */
typedef struct apt_tree apt_tree;

/* PhylogeneticTree() (:88-95).  Allocates the device pool and streams on params->device. */
int  apt_tree_create (const apt_params *params, apt_tree **out);
void apt_tree_destroy (apt_tree *t);

/* Batched addGenome (:268-302, :553-650, :657-738): the result is DEFINED as identical to calling
 * the reference's addGenome once per birth, in order.  Per birth i: payload[i*words ..],
 * compat[i*3 ..], gid[i] = genealogy().self.gid, mother/father gid and sid = genealogy().mother /
 * .father (APT_INVALID where absent).  Outputs (each may be NULL): out_sid[i] = InsertionResult.sid;
 * out_slot[i] = enveloppe slot the genome now occupies or -1 (InsertionResult.udata != nullptr iff
 * >= 0); out_replaced_gid[i] = gid of the representative it replaced, APT_INVALID when it was
 * appended or did not enter.  On error, births before the failing one stay committed
 * (as when the reference throws out of addGenome) and *n_done, when given, says how many. */
int apt_tree_add_genomes (apt_tree *t, uint32_t n, const uint32_t *payload, const float *compat,
                          const uint32_t *gid, const uint32_t *mother_gid, const uint32_t *mother_sid,
                          const uint32_t *father_gid, const uint32_t *father_sid,
                          uint32_t *out_sid, int32_t *out_slot, uint32_t *out_replaced_gid,
                          uint32_t *n_done);

/* Same, with the payloads already resident in device memory (d_payload: n*words u32, ordered after
 * `stream`); compat and the id arrays are host arrays as above.  The kernels read the payloads in
 * place; only the rows of births that change something are copied back for the host mirror. */
int apt_tree_add_genomes_dev (apt_tree *t, uint32_t n, const uint32_t *d_payload, const float *compat,
                              void *stream,
                              const uint32_t *gid, const uint32_t *mother_gid, const uint32_t *mother_sid,
                              const uint32_t *father_gid, const uint32_t *father_sid,
                              uint32_t *out_sid, int32_t *out_slot, uint32_t *out_replaced_gid,
                              uint32_t *n_done);

/* Page-locked host memory for the arrays handed to apt_tree_add_genomes (payload, compat, out_*):
 * the births' copy to the device then runs as one asynchronous DMA underneath the host-side
 * grouping, and results are written straight into out_sid / out_slot.  Ordinary (pageable) arrays
 * are accepted everywhere; large ones are staged through pinned memory by the engine's worker
 * threads.  The reference has no counterpart (its genomes never leave the host). */
int apt_host_alloc (uint64_t bytes, void **out);
/* Runs fn(ctx, lo, hi) over disjoint slices covering [0, n) on the tree's host worker threads (and the calling thread)
 * and returns when all are done: for a binding's own per-genome loops around apt_tree_add_genomes -- encoding genome
 * objects into the payload array, filling result objects.  fn must be safe to run concurrently on different slices and
 * must not call back into this tree.  Small n runs inline on the caller. */
int apt_tree_parallel_for (apt_tree *t, uint32_t n, void (*fn)(void *ctx, uint32_t lo, uint32_t hi), void *ctx);
void apt_host_free (void *p);

int apt_tree_del_genome (apt_tree *t, uint32_t sid);                                  /* :305-317 */
int apt_tree_register_candidate (apt_tree *t, uint32_t mother_sid, uint32_t father_sid, int dir); /* :320-329, :786-792; dir = +1 / -1 */
int apt_tree_set_step (apt_tree *t, uint32_t step);                                   /* :219-221 */
int apt_tree_step (apt_tree *t, uint32_t step, const uint32_t *alive_sids, uint32_t n); /* :245-262 (incl. stillborn trimming :816-860) */

/* out[8] = Stats{insertions, deletions, comparisons, branching} (:343-355), nextNodeID (:195),
 * width() (:163), step() (:190), _stillborns. */
int apt_tree_stats (apt_tree *t, uint32_t out[8]);
int apt_tree_reset_stats (apt_tree *t);                                               /* :358-360 */

/* Recorded callbacks since the last drain.  apt_tree_events copies min(cap, count) events and,
 * when `drain` is non-zero, removes them. */
uint32_t apt_tree_event_count (apt_tree *t);
uint32_t apt_tree_events (apt_tree *t, apt_event *out, uint32_t cap, int drain);

/* Tree dump with the layout of toJson (:887-941); genomes are written as
 * {"c":[mu,sigma_in,sigma_out],"g":{"f":{"g","s"},"m":{..},"s":{..}},"w":[payload words]}, user data
 * as the gid the slot object was last built from (:708-710).  Returns the length; copies at most cap. */
uint64_t apt_tree_to_json (apt_tree *t, char *out, uint64_t cap);

/* fromJson (:946-967) for a dump in the layout of apt_tree_to_json, into an EMPTY tree: rebuilds the
 * hierarchy (:908-927), re-derives the parents from the stored contributors (:960-962) and schedules
 * the whole pool for upload.  A dump built with another rsetSize is refused like at :949-954. */
int apt_tree_from_json (apt_tree *t, const char *text, uint64_t len);

/* Copy construction (:98-109, deepcopy :123-136): a new tree holding deep copies of the nodes reachable from the
 * root of `src` (children in their order), eligibilities and parents recomputed (:103).  Like the reference's, the
 * copy has no alive-species set, zeroed Stats and no recorded events. */
int apt_tree_clone (apt_tree *src, apt_tree **out);

/* Species removed by stillborn trimming (:816-860) since the last drain: copies min(cap, count) ids (and removes them
 * when `drain` is non-zero); with out == NULL returns the number pending.  For bindings that keep per-species objects. */
uint32_t apt_tree_trimmed (apt_tree *t, uint32_t *out_sids, uint32_t cap, int drain);

/* Node accessors used by the C++ binding (include/apt/phylogenetictree.hpp). */
typedef struct apt_node_info {
  uint32_t sid, parent_sid;           /* parent_sid = APT_INVALID for parentless nodes */
  uint32_t data[5];                   /* SpeciesData: first, last, count, currentlyAlive, pendingCandidates (speciesdata.hpp:15-31) */
  uint32_t fill;                      /* rset.size() */
  uint32_t n_children, n_contributors;
} apt_node_info;
int apt_tree_node_info (apt_tree *t, uint32_t sid, apt_node_info *out);               /* nodeAt(sid), :168-177 */
int apt_tree_node_children (apt_tree *t, uint32_t sid, uint32_t *out_sids, uint32_t cap);
int apt_tree_node_contributors (apt_tree *t, uint32_t sid, uint32_t *out_sid_count_elligible, uint32_t cap); /* triples */
int apt_tree_node_rep_gids (apt_tree *t, uint32_t sid, uint32_t *out_gids, uint32_t cap);
int apt_tree_node_distances (apt_tree *t, uint32_t sid, float *out_kxk);              /* K*K floats, row-major, NaN where absent */
int apt_tree_root (apt_tree *t, uint32_t *out_sid);                                   /* APT_INVALID when empty */
int apt_tree_alive_species (apt_tree *t, uint32_t *out_sids, uint32_t cap, uint32_t *n);

/* delGenome for n genomes at once (same effect as n calls of apt_tree_del_genome, in order). */
int apt_tree_del_genomes (apt_tree *t, uint32_t n, const uint32_t *sids);

/* Counters of the resolution pass since creation: out[0] births, [1] (birth,species) scores taken
 * from the device, [2] scores repaired on the host (some representatives changed in-batch),
 * [3] scores fully recomputed on the host, [4] kernel launches, [5] host-side distance evaluations,
 * [6] guard-band rechecks, [7] device pairs submitted to the scoring kernel, [8] representatives
 * those pairs read (distinct enveloppe rows per launch, summed), [9] scoring-kernel launches,
 * [10] rows uploaded by pool refreshes, [11] streams the host had to take over, [12] stream-kernel
 * launches, [13] births submitted to it, [14] enveloppe entries the device decided, [15] births rejected at top level
 * whose sub-species walk was decided on the device (family pass). */
int apt_tree_counters (apt_tree *t, uint64_t out[16]);

/* Time spent (cumulative): out[0] host prepare (s), [1] device H2D+kernels+D2H as waited for by the
 * host (s), [2] ordered resolution walk (s), [3] pool refresh (s), [4] scoring kernel alone,
 * CUDA events on the engine stream (ms), [5] stream-resolution kernel alone (ms), [6] family kernel alone (ms),
 * [7] growth of the engine's device and pinned buffers inside add_genomes calls (s; part of [0]-[3]: first-use cost). */
int apt_tree_timers (apt_tree *t, double out[8]);

/* The stream-resolution kernel as the device itself timed it: out[0] = sum over its launches of (last CTA out - first
 * CTA in) by the global nanosecond timer (ms), out[1] = launches so timed.  ([5] of apt_tree_timers is the CUDA-event
 * time of the same launches, which also contains the host's gap between recording the event and launching.) */
int apt_tree_kernel_span (apt_tree *t, double out[2]);

/* ------------------------------------------------------------------------------------------------
 * Stateless kernels (the pieces of the hot path, individually testable against the oracle).
 * A pool holds `n_species` enveloppes of up to K representatives; all pointers below are DEVICE
 * pointers when the function name ends in _dev, HOST pointers otherwise (the host variants stage
 * through device memory and run the same kernels).
 * ---------------------------------------------------------------------------------------------- */
typedef struct apt_pool apt_pool;

/* Device-resident enveloppe pool.  rep_payload: n_species*K*words u32 (slot-major, representative r
 * of species s at (s*K+r)*words); rep_compat: n_species*K*3 f32; fill: n_species u32 (k <= K). */
int  apt_pool_create (const apt_params *params, uint32_t n_species, const uint32_t *rep_payload,
                      const float *rep_compat, const uint32_t *fill, apt_pool **out);
void apt_pool_destroy (apt_pool *p);

/* a1/a2 over explicit (birth, species) pairs: out_score[p] (f32) and out_dist[p*K + r] (f32, r < fill).
 * Arithmetic of speciesMatchingScoreContinuous / Simicontinuous (:430-478). */
int apt_score_pairs (apt_pool *p, uint32_t n_births, const uint32_t *payload, const float *compat,
                     uint32_t n_pairs, const uint32_t *pair_birth, const uint32_t *pair_species,
                     float *out_score, float *out_dist);

/* Dense sweep: every birth against every species of the pool, scoring fused with the acceptance
 * test; nothing but per-birth winners leaves the chip.  out_best_species / out_best_score: species
 * with the strictly greatest score (lowest index on ties); out_first_accept: lowest species index
 * with score > 0 or APT_INVALID; out_n_accept: number of species with score > 0. */
int apt_classify_dense (apt_pool *p, uint32_t n_births, const uint32_t *payload, const float *compat,
                        uint32_t *out_best_species, float *out_best_score,
                        uint32_t *out_first_accept, uint32_t *out_n_accept);
int apt_classify_dense_dev (apt_pool *p, uint32_t n_births, const uint32_t *d_payload, const float *d_compat,
                            uint32_t *d_out_best_species, float *d_out_best_score,
                            uint32_t *d_out_first_accept, uint32_t *d_out_n_accept, void *stream);

/* Long bit-string genomes (1024..32768 bits, exact compatibility table): apt_classify_dense* run the distances on the
 * 5th-generation tensor cores -- Hamming distance as an exact int8 GEMM over genomes expanded to +-1 bytes,
 * (L - a.b) / 2, tcgen05.mma with the accumulators in tensor memory -- with the reference's arithmetic per
 * (birth, species) unchanged in the epilogue; results are bit-identical to the popcount sweep (environment
 * variable APT_DENSE_TC=0 selects that one).  The pool is expanded once, on the first sweep, and keeps the scratch
 * image of the births being swept (up to 2 GiB) between calls: sweeps of ONE pool must not overlap (issue them on one
 * stream, or order the streams); different pools are independent.
 *
 * apt_distance_matrix: the births x members distance kernel on its own (north_star (1)): out[i * n_b + j] =
 * Hamming distance of genome i of payload_a and genome j of payload_b, through the same tensor-core pipeline. */
/* 1 when apt_classify_dense* take the tensor-core path for this pool (see above), else 0. */
int apt_pool_uses_tensor_cores (apt_pool *p);
int apt_distance_matrix (const apt_params *params, uint32_t n_a, const uint32_t *payload_a, uint32_t n_b,
                         const uint32_t *payload_b, uint32_t *out);

/* Species-sharded dense sweep (north_star (5)): this rank's pool holds the species
 * [species_offset, species_offset + n_species) of `world` <= 8 shards with increasing ranges.  The sweep
 * kernel writes one 16-byte record per birth {best species (pool-wide index), score bits, first
 * accepting species, accept count} directly into EVERY rank's gather buffer: peer_records[q] (host
 * array of `world` device addresses, peer-mapped over NVLink, e.g. torch symmetric memory) is rank q's
 * buffer of world*n_births records, and this rank fills its [rank*n_births, (rank+1)*n_births) --
 * compute and exchange are one kernel, no collective call.  After a cross-rank barrier,
 * apt_combine_shards_dev reduces the `world` records of every birth to the pool-wide winners
 * (strictly greatest score, lowest species on ties; lowest accepting species; total count). */
int apt_classify_dense_shard_dev (apt_pool *p, uint32_t n_births, const uint32_t *d_payload, const float *d_compat,
                                  uint32_t species_offset, uint32_t rank, uint32_t world,
                                  const uint64_t *peer_records, void *stream);
int apt_combine_shards_dev (int device, uint32_t n_births, uint32_t world, const void *d_records,
                            uint32_t *d_out_best_species, float *d_out_best_score,
                            uint32_t *d_out_first_accept, uint32_t *d_out_n_accept, void *stream);

/* a8, enveloppe criterion (enveloppecriteria.cpp:87-133, 275-289) for n candidates: edist
 * n*K*K f32 (full symmetric matrices), gdist n*K f32 -> better (u8), than (u32), value (f32).
 * All four criteria run on the device (maxWeightedDist2Avg with the device's exp(): a decision closer than 1e-13
 * to a tie may differ from libm's). */
int apt_env_contribution (const apt_params *params, uint32_t n, const float *edist, const float *gdist,
                          uint8_t *out_better, uint32_t *out_than, float *out_value);

/* Births x births distances for in-batch entrants: out[p] = distance(payload[a[p]], payload[b[p]]). */
int apt_pairwise (const apt_params *params, uint32_t n_genomes, const uint32_t *payload,
                  uint32_t n_pairs, const uint32_t *idx_a, const uint32_t *idx_b, float *out);

/* Mate pre-filter of bailOutCrossver (src/core/crossover.h:222-248) for n_pairs (mother, father) pairs of a
 * population of n_genomes: out_dist[p] = distance(mother, father) (:234-236), out_compat[p] = the
 * compatibility the MOTHER perceives at that distance (femaleBiased, the default fcompat, :238-240).
 * Hamming distances are exact; compat uses the device exp() (<= 1 ulp from the host's). */
int apt_mate_compat (const apt_params *params, uint32_t n_genomes, const uint32_t *payload, const float *compat,
                     uint32_t n_pairs, const uint32_t *idx_mother, const uint32_t *idx_father,
                     float *out_dist, double *out_compat);

/* Pipe-throughput micro-benchmarks (roofline denominators not in MEASURED_PEAKS.json), best of 5:
 * out[0] popc32/s (XOR + POPC per op, as in a Hamming distance), out[1] LOP3/s, out[2] DFMA/s,
 * out[3] FFMA/s, chip-wide. */
int apt_microbench (int device, double out[4]);
/* out[0] = float<->double conversions/s (F2F, chip-wide): the rate the reference's `float avg += double c`
 * (phylogenetictree.hpp:465-472) is limited by when it is evaluated literally on the device. */
int apt_microbench_cvt (int device, double out[1]);

#ifdef __cplusplus
}
#endif

#endif /* APT_B200_H */
