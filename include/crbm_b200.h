/*
 * crbm_b200 — C ABI of the sm_100a (B200) kernels for the cRBM hot path of SECOMO (schulter/crbm 1.1.0).
 *
 * This is the drop-in boundary: what the reference obtains from its six compiled `theano.function`s
 * (secomo/convRBM.py:453-515) a host obtains from these entry points. Every pointer is a DEVICE pointer unless
 * the parameter name ends in `_host`; `stream` is a `cudaStream_t` passed as `void*`. No entry point allocates,
 * synchronises or throws; each returns CRBM_OK or a negative error code (text via crbm_last_error()).
 *
 * Notation (as in the reference): N sequences, L bases, K = num_motifs, M = motif_length, P = pooling,
 * Lh = L-M+1 hidden positions, Lv = Lh+M-1. Reference layouts: data (N,1,4,L) one-hot fp32; W (K,1,4,M);
 * b (1,K); c (1,4); hidden (N,K,1,Lh). All fp32, contiguous.
 *
 * PACKED SEQUENCES. The kernels read DNA as 2 bits/base (A0 C1 G2 T3, secomo/sequences.py:9-17), 16 bases per
 * little-endian uint32, one row per sequence: word 0 of a row is front padding (zero), base i lives in bits
 * [2*(i%16), 2*(i%16)+2) of word 1 + i/16, and the row is zero-padded at the end up to
 * `crbm_packed_row_words(L)` words (a multiple of 4, so rows are 16-byte aligned).
 *
 * HIDDEN BIT PLANES. Sampled hidden states of the persistent chain (convRBM.py:168-173, hidden length 200) are
 * kept as bit planes: uint32 H[strand][n][kt][jb][16] with kt = k/32, jb = j/16 (13 position blocks, so
 * crbm_fantasy_words() = 208 words per (strand, particle, k-tile)). The 16 words of a block are the warp ballots of
 * the tensor-core accumulator fragments that produce them: with k%32 = 16*mt + 8*h + g and j%16 = 8*nt + 2*q + e,
 * h[n,k,0,j] is bit 4*g+q of word ((mt*2+h)*2+nt)*2+e. crbm_hidden_bits_to_dense / _dense_to_bits convert to and from
 * the reference's dense (B,K,1,200) tensors.
 */
#ifndef CRBM_B200_H
#define CRBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRBM_OK                0
#define CRBM_ERR_ARG          -1   /* invalid argument (shape, null pointer, ...) */
#define CRBM_ERR_UNSUPPORTED  -2   /* valid for the reference but not built (a variant that was compiled out) */
#define CRBM_ERR_CUDA         -3   /* a CUDA runtime call failed */

#define CRBM_FANTASY_LEN     200   /* convRBM.py:168: hidden length of the persistent chain */
/* Largest shapes of the TUNED kernels (tensor-core and register-blocked paths). They are not limits of the library: longer motifs
 * and larger pooling factors run on the shape-generic kernels (csrc/generic.cuh), like the reference, which has no limit
 * (convRBM.py:68-108). */
#define CRBM_MAX_MOTIF_LEN    33
#define CRBM_MAX_POOLING      16

/* strand_mode of crbm_v2h_dense */
#define CRBM_STRAND_FWD 0          /* _bottomUpActivity(D, flip_motif=False), convRBM.py:238-243 */
#define CRBM_STRAND_RC  1          /* _bottomUpActivity(D, flip_motif=True) */
#define CRBM_STRAND_SUM 2          /* sum of both activities (bias twice): getHitProbs when doublestranded=False, :511-514 */
/* out_kind of crbm_v2h_dense */
#define CRBM_OUT_ACTIVITY    0
#define CRBM_OUT_PROBABILITY 1     /* _bottomUpProbability, convRBM.py:245-257 (normaliser P + sum exp) */

int         crbm_abi_version(void);
const char *crbm_last_error(void);                 /* host string, thread-local */
int         crbm_packed_row_words(int L);          /* row stride (uint32 words) of a packed sequence of L bases */
int         crbm_hidden_words(int K);              /* kw = ceil(K/32) */
int         crbm_fantasy_words(void);              /* uint32 words per (strand, particle, k-tile) of the bit planes */

/* ---- A0: sequence ingest (secomo/sequences.py:20-31,101-117) -------------------------------------------- */
/* one-hot (N,1,4,L) fp32 -> packed rows. *bad_count (device int, zeroed by the callee) receives the number of
 * positions that are not exactly one-hot. */
int crbm_pack_onehot(const float *onehot, int N, int L, uint32_t *packed, int row_words, int *bad_count, void *stream);
/* same, but *bad_count is added to instead of being reset: a host that validates lazily (once per synchronisation point,
 * not once per minibatch) keeps one running counter */
int crbm_pack_onehot_accumulate(const float *onehot, int N, int L, uint32_t *packed, int row_words, int *bad_count, void *stream);
/* base codes (N,L) uint8 in {0,1,2,3} -> packed rows; codes > 3 are counted in *bad_count. */
int crbm_pack_codes(const uint8_t *codes, int N, int L, uint32_t *packed, int row_words, int *bad_count, void *stream);
/* packed rows -> one-hot (N,1,4,L) fp32 */
int crbm_unpack_onehot(const uint32_t *packed, int row_words, int N, int L, float *onehot, void *stream);
/* synthetic i.i.d. uniform bases written directly in packed form (Philox, counter-based) */
int crbm_random_bases(uint32_t *packed, int row_words, int N, int L, uint64_t seed, void *stream);

/* ---- A1/A2/A7: visible -> hidden, dense output (convRBM.py:238-257, 507-514) ----------------------------- */
/* out (N,K,1,Lh) fp32. This is `theano_getHitProbs` (motifHitProbs) and the eager `_bottomUpActivity` /
 * `_bottomUpProbability` builders. out_kind=PROBABILITY needs Lh % P == 0 (as the reference's reshape does). */
int crbm_v2h_dense(const uint32_t *packed, int row_words, int N, int L,
                   const float *W, const float *b, int K, int M, int pooling,
                   int strand_mode, int out_kind, float *out, void *stream);
/* Same with padded output rows: row (n,k) starts at out + (n*K + k) * out_row_stride floats (out_row_stride >= Lh; the pad
 * floats hold unspecified finite values). With out 32-byte aligned and out_row_stride a multiple of 8 every store of the tcgen05 scan
 * kernel covers whole 32-byte sectors, which is what HBM write bandwidth needs (partial-sector writes cost ~1.7x:
 * scratch/store_align.cu). out_row_stride == Lh is the reference's contiguous layout. Padded rows need pooling == 1. */
int crbm_v2h_dense_strided(const uint32_t *packed, int row_words, int N, int L,
                           const float *W, const float *b, int K, int M, int pooling,
                           int strand_mode, int out_kind, float *out, long long out_row_stride, void *stream);
/* Name of the kernel crbm_v2h_dense_strided runs for this model shape and output layout (a 16-byte aligned output is assumed): the
 * choice depends on the shape (see csrc/train_tc_impl.cuh: rowdense_wanted); reported by bench.py beside the roofline numbers. */
const char *crbm_v2h_dense_kernel_name(int K, int M, int pooling, int out_kind, long long out_row_stride);

/* ---- A7 with the callers' reductions fused in (reference secomo/utils.py:155,245-247 max over positions, :306 mean) -- */
/* Per sequence and motif: maximum of the activity / hit probability over positions, the position of the first maximum
 * (numpy.argmax semantics) and the mean over positions; each output is (N,K) and may be NULL. The (N,K,1,Lh) tensor is
 * never written. pooling must be 1 for out_kind = PROBABILITY. */
int crbm_scan_reduce(const uint32_t *packed, int row_words, int N, int L,
                     const float *W, const float *b, int K, int M, int pooling,
                     int strand_mode, int out_kind, float *out_max, int *out_argmax, float *out_mean, void *stream);

/* Per-position sum over sequences of a dense chunk (reference secomo/utils.py:113-116: `motifHitProbs(seqs).mean(axis=(0,2))`, the
 * positional density): acc[k*Lh + j] += sum_{i<n} dense[(i*K + k) * row_stride + j]. `dense` is a chunk written by
 * crbm_v2h_dense_strided (row_stride = its out_row_stride); acc (K*Lh floats, zeroed by the caller before the first chunk) is
 * accumulated in sequence order, so the result does not depend on timing. */
int crbm_position_sum(const float *dense, int n, int K, int Lh, long long row_stride, float *acc, void *stream);

/* ---- A2 on a given activity tensor (eager `_bottomUpProbability(activities)`) --------------------------- */
int crbm_pool_probability(const float *act, int N, int K, int Lh, int pooling, float *prob, void *stream);

/* ---- A3: `_bottomUpSample` (convRBM.py:259-267) on a dense probability tensor ---------------------------- */
/* One uniform per pooling group, rows ordered (n,k,group) like the reference's reshape; first outcome whose
 * running sum exceeds u, none if u >= sum. `uniforms` (N*K*Lh/P floats) may be NULL -> Philox(seed, offset). */
int crbm_sample_hidden(const float *prob, int N, int K, int Lh, int pooling, const float *uniforms,
                       uint64_t seed, uint64_t offset, float *sample, void *stream);

/* ---- A5: `_topDownActivity` (convRBM.py:277-292) on dense, real-valued hidden tensors -------------------- */
/* h, hprime (N,K,1,Lh) fp32, hprime may be NULL; act (N,1,4,Lh+M-1). */
int crbm_h2v_dense(const float *h, const float *hprime, int N, int K, int Lh,
                   const float *W, const float *c, int M, float *act, void *stream);

/* ---- A6: `_topDownProbability` / `_topDownSample` (convRBM.py:294-315, 699-702) --------------------------- */
/* act (N,1,4,Lv) -> prob (4-way softmax per position; may be NULL) and/or one-hot sample (may be NULL). One
 * uniform per (n,i), n-major; `uniforms` may be NULL -> Philox(seed, offset). */
int crbm_visible_softmax_sample(const float *act, int N, int Lv, float *prob, const float *uniforms,
                                uint64_t seed, uint64_t offset, float *sample, void *stream);

/* ---- A13/A14: free energy and evaluation (convRBM.py:657-697, 466-472) ----------------------------------- */
/* fe_seq (N) = _freeEnergyForData (divided by L), fe_motif (N,K) = _freeEnergyPerMotif; either may be NULL.
 * If hsample_count != NULL (device float[N], overwritten), also draws a forward-strand hidden sample and adds,
 * per sequence, the number of active units (evaluateData's NumH numerator); uniforms from `uniforms` or Philox. */
int crbm_free_energy(const uint32_t *packed, int row_words, int N, int L,
                     const float *W, const float *b, const float *c, int K, int M, int pooling, int doublestranded,
                     float *fe_seq, float *fe_motif, float *hsample_count, const float *uniforms,
                     uint64_t seed, uint64_t offset, void *stream);

/* ---- A8-A12: one PCD/CD-k training step (convRBM.py:373-438), split where data-parallel ranks exchange ---- */
typedef struct crbm_train_args {
    /* model + optimiser state (read by accumulate, updated in place by apply) */
    float *W, *b, *c;              /* (K,1,4,M), (1,K), (1,4) */
    float *vW, *vb, *vc;           /* velocities, same shapes */
    int K, M, pooling, doublestranded, cd_k;
    float learning_rate, momentum, rho, lambda_rate;
    /* minibatch */
    const uint32_t *data;          /* packed, n_data rows of L bases (L already truncated as fit() does, :586-591) */
    int data_row_words, n_data, L;
    /* persistent chain of this rank: `batchsize` particles */
    int batchsize;
    uint32_t *fantasy_bits;        /* [2][batchsize][crbm_chain_words(K, M, pooling)]; strand 1 unused when !doublestranded */
    uint32_t *v_model;             /* packed scratch, batchsize rows of 199+M bases (holds the last sampled V) */
    int v_row_words;
    /* randomness: Philox(seed, step) unless the matching pointer is non-NULL (tests replay reference uniforms) */
    uint64_t seed, step;
    const float *u_v;              /* [cd_k][batchsize*(199+M)] or NULL */
    const float *u_h, *u_hp;       /* [cd_k][batchsize*K*(200/P)] or NULL */
    /* optional device-resident cursors so that one captured CUDA graph can be replayed for every step: if
     * step_dev != NULL the Philox step is *step_dev (seed/step fields still key the stream) and apply
     * increments it; if batch_index_dev != NULL the minibatch is rows [i*n_data, (i+1)*n_data) of `data`
     * with i = *batch_index_dev, and apply advances i modulo n_batches. */
    unsigned long long *step_dev;
    int *batch_index_dev;
    int n_batches;
    /* outputs of accumulate / inputs of apply */
    float *stats;                  /* crbm_train_stats_floats(K,M) un-normalised sums; all-reduced (sum) across ranks */
    void  *workspace;              /* crbm_train_workspace_bytes(...) bytes of scratch */
    size_t workspace_bytes;
    /* optional per-step trace of the visible bias (what a training loop that logs a metric reads back after theano_trainingFct,
     * convRBM.py:524-526): if c_trace != NULL the update kernel also stores the updated c (4 floats) into slot (step % c_trace_slots)
     * of c_trace, `step` being the Philox step of the update (args->step or *step_dev). c_trace may be page-locked host memory
     * mapped into the device's address space: the result then crosses PCIe as part of the update kernel and the host reads it
     * after its next synchronisation without a copy node in the stream. */
    float *c_trace;
    int c_trace_slots;
} crbm_train_args;

size_t crbm_train_stats_floats(int K, int M);
size_t crbm_train_workspace_bytes(int K, int M);                          /* for persistent chains of up to 512 particles */
size_t crbm_train_workspace_bytes_for(int K, int M, int batchsize);       /* any chain size (args->batchsize) */
/* positive phase on `data`, cd_k Gibbs sweeps of the persistent chain, statistics -> args->stats (overwritten). */
int crbm_train_accumulate(const crbm_train_args *args, void *stream);
/* sparsity factor, momentum update of W,b,c from (all-reduced) args->stats. */
int crbm_train_apply(const crbm_train_args *args, void *stream);
/* accumulate followed by apply, for a single rank (no exchange in between). (A fused reduce+update kernel was measured
 * and was not faster than the two small launches; see DESIGN.md.) */
int crbm_train_step(const crbm_train_args *args, void *stream);
/* The same step issued in parts (bit mask), for hosts that overlap the host->device copy of the next minibatch with the
 * chain sweeps: NEGATIVE does not read args->data. The parts of one step must be issued in this order on one stream (or on
 * streams ordered by events); REDUCE needs both NEGATIVE and POSITIVE of the same step. All kernels are launched as
 * programmatic dependents of their predecessor in the stream (CRBM_PDL=0 in the environment turns that off). */
#define CRBM_PHASE_NEGATIVE 1   /* cd_k Gibbs sweeps of the persistent chain (convRBM.py:391-408) -> partial sums */
#define CRBM_PHASE_POSITIVE 2   /* P(h|data), sparsity statistic (convRBM.py:377-388, 440-451) -> partial sums */
#define CRBM_PHASE_REDUCE   4   /* partial sums -> args->stats */
#define CRBM_PHASE_APPLY    8   /* crbm_train_apply */
int crbm_train_phases(const crbm_train_args *args, int phases, void *stream);
/* One whole step fed from HOST memory (what `theano_trainingFct(one-hot batch)` is, convRBM.py:459,524-526): launches NEGATIVE on
 * `stream`; on `copy_stream` waits for `consumed_event`, copies host_onehot ((n_data,1,4,L) fp32, pinned for an asynchronous
 * copy) into `staging` (device, same size), packs it into args->data (crbm_pack_onehot_accumulate, *bad_count running) and
 * records `ready_event`; `stream` waits for it, then POSITIVE | REDUCE | APPLY, then records `consumed_event` on `stream`.
 * Events are cudaEvent_t handles owned by the caller; `consumed_event` must have been recorded once before the first call.
 * args->batch_index_dev must be NULL. host_onehot may also be a device pointer (the copy is cudaMemcpyDefault). With p2p != NULL
 * (data parallel, see below) the minibatch is this rank's shard, args->stats points at buffer `parity`, and the last two phases
 * are crbm_train_accumulate_p2p's publishing reduction and crbm_train_apply_p2p. Returns without synchronising. */
struct crbm_p2p_args;
/* The same call without a staging copy, side stream or events: `onehot` ((n_data,1,4,L) fp32) is read by the packing kernel itself,
 * so it must be DEVICE-ACCESSIBLE — device memory, or page-locked host memory mapped into the device's address space (cudaHostAlloc /
 * cudaHostRegister under unified addressing; pass the device-side alias cudaHostGetDevicePointer returns). Every kernel of the step is
 * a programmatic dependent of its predecessor on ONE stream:
 *     h2v, v2h (chain sweeps)  ->  pack (reads `onehot` over PCIe while the sweeps run; 512-byte coalesced reads)  ->  positive phase
 *     ->  reduction  ->  update
 * The buffer behind `onehot` must stay unmodified until the step has executed. *bad_count accumulates the number of positions that
 * are not one-hot. With c_trace set (above) a whole host-fed step needs no stream operation besides its kernels. */
int crbm_train_step_hostfed(const crbm_train_args *args, const struct crbm_p2p_args *p2p, const float *onehot, int *bad_count, void *stream);
/* The host-fed step with the minibatch crossing PCIe AHEAD of the step: on `copy_stream` the call enqueues two DMA copies — host_onehot
 * ((n_data,1,4,L) fp32, page-locked) -> staging_dev (device, same size), then the 4-byte sequence value *seq_host (page-locked) ->
 * *ready_flag_dev — and on `stream` the same chain of kernels as crbm_train_step_hostfed, whose packing kernel reads staging_dev after
 * it has seen *ready_flag_dev == *seq_host (only the copy engine is waited for: it makes progress whatever the SMs do). There is no
 * event or other stream operation between the kernels of `stream`; a host that runs ahead of the device (a ring of staging buffers
 * and flags, one per step in flight) hides the transfer completely. The caller must not reuse staging_dev / ready_flag_dev / seq_host
 * before the step has executed. If the flag does not arrive within ~2 s the step proceeds and *bad_count is set to INT_MIN. */
int crbm_train_step_staged(const crbm_train_args *args, const struct crbm_p2p_args *p2p, const float *host_onehot, float *staging_dev,
                           unsigned int *ready_flag_dev, const unsigned int *seq_host, int *bad_count, void *copy_stream, void *stream);
int crbm_train_step_onehot(const crbm_train_args *args, const struct crbm_p2p_args *p2p, const float *host_onehot, float *staging,
                           int *bad_count, void *copy_stream, void *consumed_event, void *ready_event, void *stream);

/* ---- data-parallel update without a separate collective: one-shot all-reduce over NVLink peer memory, fused into the update
 * Every rank exposes, in peer-mapped memory (CUDA IPC / VMM; torch symmetric memory on the Python host), a double-buffered
 * statistics vector and a flag array [n_ranks][2]. crbm_train_accumulate_p2p is crbm_train_accumulate writing into buffer
 * `parity` (args->stats must point at it). crbm_train_apply_p2p first stores `sequence` into entry [rank][parity] of EVERY rank's
 * flag array (one-way NVLink stores after a system-scope fence: this rank's statistics are complete, the reduction being the
 * predecessor grid), then polls its own flag array until all ranks carry `sequence`, then every thread loads its statistics
 * entries from all ranks' buffers at once (one NVLink round trip), adds them in rank order (bit-identical sums on every rank, so
 * the replicas stay identical) and applies the update with the arithmetic of crbm_train_apply: no NCCL launch, no extra kernel,
 * and whole steps sit in one CUDA graph. With n_ranks == 1 (the rank as its own peer, ordinary device memory) the result is
 * bit-identical to crbm_train_step.
 * The wait is bounded (~2 s); on expiry *error_flag is set to 1. `sequence` must increase by one per step on every rank and
 * consecutive steps must alternate `parity`; when args->step_dev is set the sequence is *step_dev + 1. */
#define CRBM_MAX_PEERS 8
typedef struct crbm_p2p_args {
    int n_ranks, rank, parity;
    unsigned long long sequence;
    float *peer_stats[CRBM_MAX_PEERS];               /* each rank's stats double buffer: 2 * stats_stride floats */
    unsigned long long *peer_flags[CRBM_MAX_PEERS];  /* each rank's flag array: n_ranks * 2 entries, zero-initialised */
    size_t stats_stride;                             /* floats between the two buffers (>= crbm_train_stats_floats) */
    int *error_flag;                                 /* device int, set to 1 if a peer never showed up */
    unsigned int *ticket;                            /* device counter (zero-initialised) electing the last CTA */
    /* optional receive buffers of the FUSED exchange (crbm_train_step_p2p), peer-mapped, zero-initialised, per rank
     * 2 * ll_parity_stride 8-byte elements with ll_parity_stride >= crbm_p2p_ll_elems(K, M, pooling); NULL: not used */
    unsigned long long *peer_ll[CRBM_MAX_PEERS];
    size_t ll_parity_stride;
} crbm_p2p_args;
int crbm_train_accumulate_p2p(const crbm_train_args *args, const crbm_p2p_args *p2p, void *stream);
/* One whole data-parallel step. Where the model shape trains on the mma.sync kernels (small un-pooled banks such as C2), every rank
 * has rows in the step and p2p->peer_ll is set, reduction, all-reduce and update are ONE launch (reduce_update_ll_kernel): a CTA
 * that has reduced its rows of the per-CTA partials pushes them into every rank's receive buffer as 8-byte (value, sequence number)
 * stores — an aligned 8-byte store is one transaction, so the sequence number proves its value has arrived: no fence, no flag —
 * polls its own buffer, adds the ranks in rank order (bit-identical on every rank) and applies the update; one NVLink one-way latency
 * per step. Otherwise crbm_train_accumulate_p2p + crbm_train_apply_p2p. All ranks must take the same branch: pass peer_ll only
 * when no rank's shard of the step is empty. crbm_p2p_ll_elems returns 0 for shapes the fused exchange does not serve. */
int crbm_train_step_p2p(const crbm_train_args *args, const crbm_p2p_args *p2p, void *stream);
size_t crbm_p2p_ll_elems(int K, int M, int pooling);
int crbm_train_apply_p2p(const crbm_train_args *args, const crbm_p2p_args *p2p, void *stream);

/* layout of `stats` (offsets in floats), for hosts that want to inspect gradients */
typedef struct crbm_stats_layout {
    int vh_data, vh_model, vh_q;   /* each K*4*M: sum P(x)V folded over strands; same for the model; sum Q(x)V */
    int h_data, h_model, h_q, h_fwd; /* each K: sum P+P' (data), same (model), sum Q, sum P forward only */
    int v_data, v_model;           /* each 4: base counts */
    int n_hidden_data, n_hidden_model, n_vis_data, n_vis_model; /* scalars: N*Lh, B*200, N*L, B*(199+M) */
    int total;
} crbm_stats_layout;
int crbm_train_stats_layout(int K, int M, crbm_stats_layout *out_host);

/* fantasy particles <-> the reference's dense (B,K,1,200) fp32 tensors (tests, inspection): fragment-ordered bit planes */
int crbm_hidden_bits_to_dense(const uint32_t *bits, int B, int K, float *dense, void *stream);
int crbm_hidden_dense_to_bits(const float *dense, int B, int K, uint32_t *bits, void *stream);

/* THE PERSISTENT CHAIN OF A MODEL (convRBM.py:168-173). Its device layout follows the kernels that train the model shape:
 *   layout 0  fragment-ordered bit planes (above; the mma.sync kernels: motif_length > 32 or pooling not in {1,2,4,5,10})
 *   layout 1  linear planes of the tcgen05 kernels: uint32 H[strand][n][w][Kpad], Kpad = 64 ceil(K/64), VP = (32/P) P;
 *             bit b (< VP) of word w of motif k is hidden position w * VP + b
 * args->fantasy_bits of the training entry points is uint32 [2][batchsize][crbm_chain_words(K, M, pooling)]; one strand plane
 * (batchsize * crbm_chain_words words) converts to / from the reference's dense tensor with the two functions below. */
int    crbm_chain_layout(int K, int M, int pooling);
size_t crbm_chain_words(int K, int M, int pooling);
int    crbm_chain_to_dense(const uint32_t *bits, int B, int K, int M, int pooling, float *dense, void *stream);
int    crbm_chain_from_dense(const float *dense, int B, int K, int M, int pooling, uint32_t *bits, void *stream);

/* ---- A0, host side of the streaming ingest (secomo/sequences.py:33-52 readSeqsFromFasta + :20-31 the base table) ------------
 * One pass over `n` bytes of COMPLETE multi-FASTA records (HOST pointers): header lines dropped, line ends / blanks inside a
 * record dropped, letters mapped A0 C1 G2 T3 case-insensitively into codes[record][L]; records containing N/n are skipped and
 * counted in *n_skipped (the reference prints "skip sequence containing N" for each). *L_io is the record length (0 = not yet
 * known: taken from the first record). Returns CRBM_OK, -11 if a record's length differs (the reference's one-hot array cannot be
 * built), -12 for any other letter (*bad_char = the letter; the reference raises KeyError). As in the reference, a record
 * containing N is skipped before its other letters or its length matter, and the record length is that of the first record
 * without N; records without any sequence character are dropped; the first defect in file order is the one reported. */
int crbm_fasta_parse_block_host(const unsigned char *buf_host, size_t n, unsigned char *codes_host, size_t capacity_records,
                                int *L_io, long long *n_records, long long *n_skipped, int *bad_char);

/* The same pass, multi-threaded, writing rows of 2-bit bases in the device layout (row_words >= crbm_packed_row_words(L) words per
 * record, pad words zeroed) — into page-locked memory, one cudaMemcpyAsync away from a PackedSequences tensor: readSeqsFromFasta +
 * seqToOneHot (secomo/sequences.py:33-52, 101-117) without the one-hot array and without a packing kernel. `n_threads` <= 0: one per
 * host core (at most one per MiB of input). Returns CRBM_OK, -11 / -12 as above, or -13 if `rows_host` is NULL or holds fewer than the
 * *n_records rows the block may need (an upper bound: its number of headers; *L_io is set) — size the buffer and call again. */
int crbm_fasta_pack_block_host(const unsigned char *buf_host, size_t n, uint32_t *rows_host, int row_words, size_t capacity_rows,
                               int *L_io, long long *n_records, long long *n_skipped, int *bad_char, int n_threads);

/* A0 on the host threads: one-hot float32 (N,1,4,L) in HOST memory (sequence / plane strides in floats, positions contiguous) ->
 * packed rows in HOST memory (page-locked: one cudaMemcpyAsync of 0.25 byte per base follows), with the validity rule of
 * crbm_pack_onehot; *bad_count = number of positions that are not one-hot. The reference-facing calls take NumPy one-hot arrays
 * (secomo/sequences.py:101-117 builds them, 16 bytes per base): packing them where they lie saves 63/64 of the PCIe traffic. */
int crbm_pack_onehot_host(const float *onehot_host, long long N, int L, long long seq_stride, long long plane_stride,
                          uint32_t *rows_host, int row_words, long long *bad_count, int n_threads);

/* launch counter: number of kernels this library has launched since load (bench.py's `gpu_launches`) */
unsigned long long crbm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CRBM_B200_H */
