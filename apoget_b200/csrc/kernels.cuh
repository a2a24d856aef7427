// Device-side data views and kernel launchers of libapt_b200.so (sm_100a only).
#ifndef APT_B200_KERNELS_CUH
#define APT_B200_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace apt {

constexpr uint32_t KMAX = 32;        // largest supported enveloppe size (rsetSize)
constexpr uint32_t ORDER_INDEX_MASK = 0x3FFFFFFFu;   // StreamArgs::order: birth index bits
constexpr uint32_t ORDER_CLASS_SHIFT = 30;

/// Enveloppe pool in HBM.  Representative r of the species in slot s lives at
/// payload[(s*K + r) * wp .. +wp) (wp = payload words padded to a multiple of 4 so every row is
/// uint4-aligned; padding words are zero and do not change either distance).
/// cmeta[s*K + r] = {mu, sigma_in, sigma_out, bits(class id)}.
struct PoolView {
  const uint32_t *payload;
  const float4 *cmeta;
  const uint32_t *fill;       // representatives present per slot (k <= K)
  uint32_t K, wp, n_slots;
};

/// Scoring configuration (config::PTree thresholds + the compatibility table).
/// With use_lut, lut[cls * lut_stride + d] holds compatibility(d) of class cls as computed by the
/// HOST libm for every integer distance d in [0, 32*words] -- the device then reproduces the
/// reference's doubles bit for bit.
struct ScoreCfg {
  int kind;                   // APT_BITS / APT_F32
  int full_continuous;
  int use_lut;
  double compat_thr, sim_thr, avg_thr;
  const double *lut;
  uint32_t lut_stride;
  uint32_t lut_classes;       // rows of the table
};

/// K1: one warp per (birth, species) pair.
void launch_score_pairs (const PoolView &pool, const ScoreCfg &cfg,
                         const uint32_t *b_payload, const float4 *b_cmeta,
                         uint32_t n_pairs, const uint32_t *pair_birth, const uint32_t *pair_slot,
                         float *out_score, float *out_dist, cudaStream_t stream);

/// Pads n rows of `words` u32 into rows of `wp` u32 and packs compat triples + class ids into
/// float4 metadata (device-resident inputs of the *_dev entry points).
void launch_pack_births (const uint32_t *src_payload, const float *src_compat, const uint32_t *cls,
                         uint32_t n, uint32_t words, uint32_t wp,
                         uint32_t *dst_payload, float4 *dst_cmeta, cudaStream_t stream);

/// dst[0..n) = value
void launch_fill_f4 (float4 *dst, uint32_t n, float4 value, cudaStream_t stream);

/// dst[i][0..wp) = src[i][0..words) padded with zeros
void launch_pad_rows (const uint32_t *src, uint32_t n, uint32_t words, uint32_t wp, uint32_t *dst, cudaStream_t stream);

/// dst[i][0..words) = src[idx[i]][0..words)
void launch_gather_rows (const uint32_t *src, const uint32_t *idx, uint32_t n, uint32_t words, uint32_t *dst,
                         cudaStream_t stream);

/// dst[e][0..words) = src[list[2e]][0..words) for e < min(*count, max_entries), `count` being a device counter
/// The births of a pass that are not settled on the device, in arrival order: list[2 e] = birth, list[2 e + 1] = its result
/// word; *total = their number.  block_count: scratch of list_unsettled_blocks(n) words.  fam_act: the family pass'
/// verdicts (births that joined a sub-species without changing it are settled too) or null.
uint32_t list_unsettled_blocks (uint32_t n);
void launch_list_unsettled (const uint32_t *act, const uint32_t *fam_act, uint32_t n, uint32_t *block_count,
                            uint32_t *list, uint32_t *total, cudaStream_t stream);
/// The births of a pass grouped into per-species streams on the device (what Engine::groupParallel does on the host), one
/// cooperative launch.  msid / fsid: device-accessible (page-locked host memory is read in place).  key[n], start / cursor /
/// gidx [ns], rest[n]: device scratch.  total[ns], info[8]: this call's counters, zero on entry; the kernel clears
/// total_next / info_next, the sets the NEXT call passes as total / info.  order[n] (device): the streams, species in
/// ascending id, births in arrival order, as the stream kernel reads them.  Written to page-locked host memory directly:
/// h_total[sid] births per species, h_info = {streams, longest, births in streams, 1 when a stream was too long to be put
/// back into arrival order here (nothing below is valid then), births outside the streams}, h_rest those births (in no
/// particular order), h_order a copy of order, h_birth_group[i] the birth's stream (rank of its species among the
/// species with births) or -1.
void launch_group_births (const uint32_t *msid, const uint32_t *fsid, uint32_t n, const uint32_t *sid_slot, uint32_t ns,
                          uint32_t root_sid, uint32_t *key, uint32_t *total, uint32_t *total_next, uint32_t *start, uint32_t *cursor,
                          uint32_t *gidx, uint32_t *info, uint32_t *info_next, uint32_t *order, uint32_t *rest, uint32_t *h_total,
                          uint32_t *h_info, uint32_t *h_rest, uint32_t *h_order, int32_t *h_birth_group, cudaStream_t stream);
void launch_gather_listed_rows (const uint32_t *src, const uint32_t *list, const uint32_t *count, uint32_t max_entries,
                                uint32_t words, uint32_t *dst, cudaStream_t stream);

/// Scatters staged enveloppe rows into the pool: row i of the staging buffers goes to pool row
/// dst_row[i] (= slot*K + r).
void launch_scatter_rows (const uint32_t *stage_payload, const float4 *stage_cmeta, const uint32_t *dst_row,
                          uint32_t n_rows, uint32_t wp, uint32_t *pool_payload, float4 *pool_cmeta,
                          cudaStream_t stream);

/// dst[idx[i]] = val[i]
void launch_scatter_u32 (const uint32_t *idx, const uint32_t *val, uint32_t n, uint32_t *dst, cudaStream_t stream);

/// a8: enveloppe criterion, one thread per candidate (crit 0 = maxAverage, 1 = maxMinDist).
void launch_env_contribution (int crit, uint32_t K, uint32_t n, const float *edist, const float *gdist,
                              uint8_t *out_better, uint32_t *out_than, float *out_value, cudaStream_t stream);

/// Births x births distances, one warp per pair.
void launch_pairwise (int kind, uint32_t wp, const uint32_t *payload, uint32_t n_pairs,
                      const uint32_t *idx_a, const uint32_t *idx_b, float *out, cudaStream_t stream);

/// Per-birth result word of the stream kernel
enum : uint32_t {
  ACT_UNPROCESSED = 0u,       ///< the warp stopped before this birth (host takes over)
  ACT_FALLTHROUGH = 1u,       ///< score <= 0 at top level: the host walks the sub-species
  ACT_ACCEPTED = 2u,          ///< accepted, enveloppe unchanged
  ACT_ENTERED = 3u,           ///< accepted and now a representative
  ACT_CODE_MASK = 3u,
  ACT_REPLACED = 4u,          ///< (with ACT_ENTERED) took the place of an older representative
  ACT_UNCERTAIN = 8u,         ///< decision needs exact arithmetic (guard band): host takes over from here
  ACT_SLOT_SHIFT = 4u,        ///< 5 bits: enveloppe slot
  ACT_ROW_SHIFT = 9u          ///< 23 bits: index of the birth's distance row in `rows`
};

/// Arguments of resolve_streams_kernel (stream.cu): one warp per species stream.
struct StreamArgs {
  PoolView pool;              ///< state at submission; the kernel writes the final state back through:
  uint32_t *pool_payload_w;  float4 *pool_cmeta_w;  uint32_t *pool_fill_w;
  float *pool_dist;           ///< [slot][K][K] intra-enveloppe distances (DistanceMap, treetypes.h:259)
  ScoreCfg cfg;
  int crit;                   ///< 0 maxAverage, 1 maxMinDist
  float guard;                ///< |score| below which a non-table score is not trusted
  const uint32_t *b_payload;  const float4 *b_cmeta;
  uint32_t n_groups;
  uint32_t max_stream;        ///< births in the longest stream (picks the CTA width)
  const uint32_t *grp_slot;   ///< [n_groups] pool slot of the stream's species
  const uint32_t *grp_off;    ///< [n_groups+1] offsets into order
  const uint32_t *order;      ///< births grouped by stream, arrival order inside a stream: low 30 bits = birth
                              ///< index, high 2 bits = what the birth contributes to its own species
                              ///< (0 primordial: nothing, 1 one parent, 2 both parents; :576-577)
  const uint32_t *grp_sid;    ///< [n_groups] species id of the stream's species
  uint32_t *out_sid;          ///< [n_births] InsertionResult.sid of accepted births (pre-set to APT_INVALID)
  int32_t *out_slot;          ///< [n_births] enveloppe slot of entered births (pre-set to -1)
  uint4 *grp_sum;             ///< [n_groups] {accepted-unchanged births contributing 0, 1, 2; stopped early}
  uint32_t *act;              ///< [n_births] result words (zero-initialised by the caller)
  float *rows;                ///< [n_rows][K] cached distances of the births that entered an enveloppe
  uint32_t *row_payload;      ///< [n_rows][wp] their genomes (device-resident births: what the host mirrors), or null
  uint32_t *ft_list;          ///< births rejected at top level, appended as (birth, stream) pairs ...
  uint32_t *ft_counter;       ///< ... counted here (zero-initialised)
  uint32_t *row_counter;      ///< zero-initialised; a stream claims the rows of a step (its enveloppe entries) with one add
  unsigned long long *span = nullptr;   ///< optional, zero-initialised: [0] max over the CTAs of ~(global timer at entry), [1] max of the timer at exit
  uint32_t *rej_order = nullptr;  ///< optional [n stream births]: per stream, at grp_off[g], its births rejected at top level in ARRIVAL
                              ///< order (order words) -- the input of the sub-species resolution on the device (family.cu)
  uint32_t *rej_fill = nullptr;   ///< same layout: the enveloppe fill each one's top-level score saw (Stats::comparisons)
  uint32_t *rej_count = nullptr;  ///< [n_groups] how many
  int lut_in_smem = 0;        ///< set by the launcher
  int preload = 0;            ///< the launcher only makes the runtime load the kernel variant these arguments select (no launch)
  cudaEvent_t ev_before = nullptr;   ///< host side: recorded on the stream by the launcher immediately in front of the launch
  unsigned long long *prof = nullptr;   ///< optional [n_groups][12] per-stream timeline (APT_STREAM_PROF), see stream.cu
};
bool stream_kernel_supports (uint32_t wp, uint32_t K, int crit);
void launch_resolve_streams (const StreamArgs &a, cudaStream_t stream);

/// Result word of the family kernel (family.cu) per birth rejected at top level; zero = not processed (the host walks it)
enum : uint32_t {
  FAM_UNPROCESSED = 0u,
  FAM_JOINED = 1u,            ///< accepted by a sub-species whose enveloppe it leaves alone
  FAM_ENTERED = 2u,           ///< accepted by a sub-species and now one of its representatives
  FAM_FOUNDED = 3u,           ///< no sub-species accepted it: founds a new one (:637-643)
  FAM_CODE_MASK = 3u,
  FAM_REPLACED = 4u,          ///< (with FAM_ENTERED) took the place of an older representative
  FAM_SLOT_SHIFT = 3u,        ///< 5 bits: enveloppe slot
  FAM_CHILD_SHIFT = 8u        ///< 24 bits: the sub-species: rank in the family's snapshot list (visiting order), or
                              ///< snapshot count + j for the j-th species founded in this pass
};
/// out_sid of a birth placed in a species founded in this pass: FAM_NEW_SID | family * max_new + j (the id is given out
/// by the host's ordered commit, :225-230)
constexpr uint32_t FAM_NEW_SID = 0x80000000u;

/// K3 -- findBestDerived + insertInto + makeNode (phylogenetictree.hpp:493-548, :657-726, :637-643) on the device for
/// whole "families": the births one stream's species rejected at top level, IN ORDER, against that species' direct
/// sub-species -- those it had at submission and those founded inside this very pass -- with every enveloppe change
/// (append, replacement, a new species) visible to the next birth, exactly as the reference's sequential walk.
/// One CTA per family; the sub-species' enveloppes live in shared memory.  A family is only submitted when nothing
/// else in the pass can touch its sub-species (engine.cpp: no stream or two-parent birth of their own).
struct FamilyArgs {
  PoolView pool;
  uint32_t *pool_payload_w;  float4 *pool_cmeta_w;  uint32_t *pool_fill_w;  float *pool_dist;
  ScoreCfg cfg;
  int crit;                   ///< 0 maxAverage, 1 maxMinDist
  float guard;
  const uint32_t *b_payload;  const float4 *b_cmeta;
  uint32_t n_fam;
  const uint32_t *fam_group;  ///< [n_fam] stream index of the family
  const uint32_t *grp_off;    ///< [n_groups+1] stream extents = where its ordered reject list starts in rej_order
  const uint32_t *rej_order, *rej_fill, *rej_count;     ///< written by the stream kernel (StreamArgs)
  const uint32_t *child_off;  ///< [n_groups+1] offsets into child_slot / child_sid
  const uint32_t *child_slot; ///< sub-species at submission, visiting order (newest first, :506): pool slots ...
  const uint32_t *child_sid;  ///< ... and species ids
  uint32_t *slot_counter;     ///< zero-initialised: species founded in this pass take the pool slots slot_base + 0, 1, 2 ...
  uint32_t slot_base, slot_limit;   ///< ... up to slot_limit of them (the pool arrays are sized for it)
  uint32_t max_new;           ///< species a family may found before it hands the rest of its births back
  uint32_t cap;               ///< sub-species (at submission + founded) that fit in shared memory
  uint32_t *fam_act;          ///< [n_births] result words (zero-initialised)
  uint32_t *fam_row;          ///< [n_births] index into rows / row_payload (FAM_ENTERED, FAM_FOUNDED)
  uint32_t *out_sid;  int32_t *out_slot;
  float *rows;                ///< [n_rows][K] cached distances of the births that entered an enveloppe
  uint32_t *row_payload;      ///< [n_rows][wp] their genomes, or null
  uint32_t *row_counter;      ///< zero-initialised
  uint4 *fam_info;            ///< [n_fam] {births decided, species founded, 1 = handed the rest back, sub-species in the end}
  unsigned long long *fam_stats;  ///< [n_fam][2] Stats::branching (:526) and Stats::comparisons (:444 / :470) of the births
                              ///< decided here (their top-level scores included)
  const uint32_t *fam_join_off;   ///< [n_fam] where the family's per-sub-species counters start in child_join
  uint32_t *child_join;       ///< per sub-species (snapshot ranks, then founded): FAM_JOINED births contributing 1 / 2 (:576-577)
  uint32_t *new_founder;      ///< [n_fam][max_new] birth that founded the j-th new species
  uint32_t *new_slot;         ///< [n_fam][max_new] the pool slot it lives in
  int lut_in_smem = 0;
  int preload = 0;            ///< as StreamArgs::preload
};
bool family_kernel_supports (uint32_t wp, uint32_t K, int crit);
/// sub-species capacity of one family for this shape (0 = unsupported)
uint32_t family_capacity (uint32_t wp, uint32_t K, const ScoreCfg &cfg);
void launch_resolve_families (const FamilyArgs &a, cudaStream_t stream);

/// K2 -- findBestDerived (phylogenetictree.hpp:493-548) for the births a stream rejected at top level:
/// one warp per rejected birth walks the sub-species of its species in the reference's visiting
/// order (newest first) and stops at the first score > 0 (:540-541).
struct ChildArgs {
  PoolView pool;
  ScoreCfg cfg;
  float guard;
  const uint32_t *b_payload;  const float4 *b_cmeta;
  const uint32_t *ft_list;    ///< (birth, stream) pairs written by the stream kernel
  const uint32_t *ft_counter; ///< how many
  uint32_t max_ft;            ///< capacity (grid is sized for it; surplus warps exit)
  const uint32_t *child_off;  ///< [n_groups+1] offsets into child_slot
  const uint32_t *child_slot; ///< sub-species slots per stream, visiting order; APT_INVALID = not on the device
  uint32_t *out_first;        ///< [max_ft] rank of the first accepting sub-species, or APT_INVALID
  uint32_t *out_flags;        ///< [max_ft] bit 0: the walk stopped at out_first because its score needs the host
  float *out_score;           ///< [max_ft] score at out_first
  float *out_dist;            ///< [max_ft][K] distances to the representatives of out_first
};
void launch_classify_children (const ChildArgs &a, cudaStream_t stream);

/// dst[idx[i]*len .. +len) = src[i*len .. +len)
void launch_scatter_blocks_f32 (const float *src, const uint32_t *idx, uint32_t n, uint32_t len, float *dst,
                                cudaStream_t stream);

/// Mate pre-filter (bailOutCrossver, src/core/crossover.h:222-248): per (mother, father) pair the genomic
/// distance and the compatibility the MOTHER perceives at that distance (femaleBiased default).
void launch_mate_compat (int kind, uint32_t wp, const uint32_t *payload, const float4 *cmeta, uint32_t n_pairs,
                         const uint32_t *idx_mother, const uint32_t *idx_father, float *out_dist, double *out_compat,
                         cudaStream_t stream);

/// Dense sweep (every birth x every species), fused score / threshold / argmax.
void launch_classify_dense (const PoolView &pool, const ScoreCfg &cfg,
                            const uint32_t *b_payload, const float4 *b_cmeta, uint32_t n_births,
                            uint32_t *out_best_species, float *out_best_score,
                            uint32_t *out_first_accept, uint32_t *out_n_accept, cudaStream_t stream);

/// Species-sharded dense sweep: where the per-birth winners of a shard go.  With
/// world > 0 the sweep writes ONE 16-byte record per birth {best species (pool-wide index), score
/// bits, first accepting species, accept count} straight into every peer's gather buffer at
/// peer[p][rank * n_births + birth] -- peer-mapped memory, i.e. NVLink stores from inside the sweep
/// kernel -- instead of the four local arrays; the exchange needs no separate collective.
constexpr uint32_t MAX_PEERS = 8;
struct ShardSink {
  uint4 *peer[MAX_PEERS];
  uint32_t world = 0, rank = 0, species_offset = 0;
};

/// Register-tiled dense sweep (dense.cu); launch_classify_dense picks it whenever it fits.
bool dense_tiled_supports (uint32_t wp, uint32_t K);
void launch_classify_dense_tiled (const PoolView &pool, const ScoreCfg &cfg, const uint32_t *b_payload,
                                  const float4 *b_cmeta, uint32_t n_births, uint32_t *out_best_species,
                                  float *out_best_score, uint32_t *out_first_accept, uint32_t *out_n_accept,
                                  cudaStream_t stream, const ShardSink *sink = nullptr);

/// Tensor-core dense sweep for long bit-string genomes (tc_dense.cu): Hamming distance as an exact int8 GEMM
/// (tcgen05.mma, accumulators in TMEM) over operands pre-expanded into tile images, same epilogue arithmetic.
bool dense_tc_supports (const ScoreCfg &cfg, uint32_t wp, uint32_t K);
size_t tc_pool_image_bytes (uint32_t n_species, uint32_t K, uint32_t wp);
size_t tc_pool_meta_bytes (uint32_t n_species, uint32_t K);
size_t tc_births_image_bytes (uint32_t n_births, uint32_t wp);
void launch_tc_expand_pool (const PoolView &pool, uint8_t *img, uint32_t *meta, cudaStream_t stream);
void launch_tc_expand_births (const uint32_t *b_payload, uint32_t n_births, uint32_t wp, uint8_t *img, cudaStream_t stream);
/// out_matrix != null: distance-matrix mode, out_matrix[b * ld + row] = Hamming distance (ld = tiles * 256 pool rows)
void launch_classify_dense_tc (const PoolView &pool, const ScoreCfg &cfg, const uint8_t *a_img, const uint8_t *b_img,
                               const uint32_t *tile_meta, const float4 *b_cmeta, uint32_t n_births,
                               uint32_t *out_best_species, float *out_best_score, uint32_t *out_first_accept,
                               uint32_t *out_n_accept, uint32_t *out_matrix, cudaStream_t stream,
                               const ShardSink *sink = nullptr, uint32_t rec_stride = 0, uint32_t rec_offset = 0);

/// Local winners -> records in every peer's gather buffer (the sweep kernels that do not write
/// records themselves: genome shapes the tiled kernel does not cover)
void launch_publish_records (const uint32_t *best, const float *best_score, const uint32_t *first, const uint32_t *count,
                             uint32_t n_births, const ShardSink &sink, cudaStream_t stream);

/// Pool-wide winners from the gathered records rec[rank * n_births + birth] of `world` shards holding
/// increasing species ranges: strictly greatest score (lowest species on ties = first shard holding
/// the maximum), lowest accepting species, total accept count.
void launch_combine_shards (const uint4 *rec, uint32_t n_births, uint32_t world, uint32_t *out_best_species,
                            float *out_best_score, uint32_t *out_first_accept, uint32_t *out_n_accept,
                            cudaStream_t stream);

} // namespace apt

#endif
