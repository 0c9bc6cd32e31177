// ust_dev.h — structures shared by the host side of libust.so (ust_api.cu) and its kernels.
#pragma once
#include <stdint.h>

#include "../../include/ust.h"
#include "ust_lut.h"

#define UST_THREADS 256          /* auxiliary kernels */
#ifndef UST_VERIFY_THREADS
#define UST_VERIFY_THREADS 384   /* verification kernel */
#endif
#define UST_MAX_CTAS 1024        /* per-CTA diagnostic stamps */
#define UST_MAX_WORLD 8
#define UST_DS_SMEM_MAX 1024

// Streaming kernel geometry (ust_stream.cu): a tile is the unit a CTA claims, the TMA engine copies into one ring
// stage, and the slot speculation is made for. Tiles of fewer nodes (a multiple of 128) are used for small
// snapshots so that every SM gets work; the ring stages are sized for the largest.
#ifndef UST_TILE_NODES
#define UST_TILE_NODES 3072
#endif
#ifndef UST_STAGES
#define UST_STAGES 4
#endif
#ifndef UST_CONSUMER_WARPS
#define UST_CONSUMER_WARPS 12
#endif
#define UST_STREAM_THREADS (32 * (1 + UST_CONSUMER_WARPS))

// Exchange vector (int64 lanes): what one shard contributes to / learns from the cluster-wide
// constraint arithmetic (upgrade_inplace.go:49-62). Summed across shards; per-rank slots are one-hot,
// so the sum doubles as an all-gather.
#define UST_V_HIST 0                         /* 16 lanes: nodes per state code */
#define UST_V_UNAVAILABLE 16
#define UST_V_CANDIDATES 17
#define UST_V_RANK_CAND (18)                 /* UST_MAX_WORLD lanes */
#define UST_V_RANK_NODES (18 + UST_MAX_WORLD)
#define UST_V_RANK_ERRINV (18 + 2 * UST_MAX_WORLD) /* ~abort key of the rank, 0 = none */
#define UST_V_LEN (18 + 3 * UST_MAX_WORLD)

// Peer mailboxes of the fused multi-GPU exchange: rank s writes its lanes into slot[epoch & 1][s] of EVERY rank's
// mailbox over NVLink (CUDA IPC mapping). Every 8-byte word carries half a lane and the call number
// ({data:32, epoch:32}, one 8-byte store each - delivered whole), so the data is its own flag: no fence, no separate
// flag store, the reader polls the words it needs.
#define UST_MBOX_WORDS (2 * UST_V_LEN)   /* 84 words = 672 B per slot */
struct UstMailbox {
  unsigned long long slot[2][UST_MAX_WORLD][UST_MBOX_WORDS + 12];
};

// Device workspace owned by a handle. The per-call accumulators exist twice: call k uses set (k & 1); the streaming
// kernel of call k+1 clears the set of call k (the verification kernel of call k, the only reader, has completed by
// then - stream order), so nobody ever waits for a reset. Invariant between calls: the set of the NEXT call is zero.
#define UST_MAX_SEGMENTS 16   /* streaming launches per call (pipelined uploads): one ticket each */
struct UstWorkspace {
  unsigned long long acc[2][18];  // hist[0..13], -, -, unavailable, candidates (this shard)
  unsigned long long errinv[2];   // ~min abort key seen while streaming, 0 = none
  unsigned int ticket[2][UST_MAX_SEGMENTS];  // dynamic tile claiming, one counter per streaming launch of the call
  int spec_used[2];               // the speculative cut the streaming kernel ran with
  unsigned long long bs_acc[18];  // BuildState kernels (zero between calls: their finish kernel clears it)
  unsigned int arrive;            // split mode: CTAs of the publishing streaming launch that have finished
  unsigned int comm_timeout;      // set when a peer did not show up (kernel gives up instead of hanging)
  // fused exchange: CTA 0 of the verification kernel talks to the peers; it leaves the summed vector here for the
  // other CTAs of its grid and then sets xflag = (epoch << 32) | ok (by epoch parity, like the mailboxes)
  long long xsum[2][UST_V_LEN];
  unsigned long long xflag[2];
  // Speculation hint carried from call to call (results never depend on it, only how many tiles are redone):
  // an earlier call's cut tile, valid for calls with the same signature (size, tiling, slot policy). One slot per
  // call parity: the verification kernel writes its own call's slot; a streaming kernel reads the previous call's slot
  // - or, when it runs beside the previous call's verification kernel (UstParams::relaxed), the slot of the call
  // before that, which nobody is writing.
  unsigned long long hint_sig[2];
  int hint_cut[2];
  unsigned long long dbg[UST_MAX_CTAS][4];  // %globaltimer stamps per streaming CTA: entry, first tile landed, stream end, exit
  unsigned long long dbg2[16];              // verification kernel, CTA 0: woken, vector loaded, decided, done; 4..: inside the decision
};

// abort key: (pass << 56) | (global node index + 1); policy-level aborts use index part 0
#define UST_KEY(pass, gidx_plus1) ((((unsigned long long)(pass)) << 56) | (unsigned long long)(gidx_plus1))

struct UstParams {
  long long n;  // nodes in this shard
  const uint8_t* hot;
  const uint32_t* flags;
  const int32_t* pod_rev;
  const int32_t* ds_idx;
  const int32_t* ds_rev;
  int n_ds;
  const int32_t* pod_off;     // nullable
  const uint16_t* pod_flags;  // nullable
  uint8_t* next;
  uint16_t* actions;
  uint8_t* outcome;  // nullable
  const uint32_t* lut;    // UST_LUT_ENTRIES words, then 16 uint2 lookup constants (ust_lut.h)
  const uint8_t* podlut;  // UST_PODLUT_ENTRIES bytes
  uint8_t* podsum;        // per-node pod-list summary (written by the pod-summary kernel, read by the streaming pass); null = no pod lists
  UstWorkspace* ws;
  unsigned int* cand_tile;  // upgrade candidates per tile (streaming pass writes, the decision reads)
  long long* xchg;        // UST_V_LEN lanes (split mode: the streaming kernel writes, the verification kernel reads the reduced copy)
  ust_counters* out;      // device
  // policy (flattened; see include/ust.h)
  long long max_parallel;
  long long max_unav_value;
  int max_unav_kind;
  int active;             // policy != nil && AutoUpgrade (upgrade_state.go:179-182); 0 => every node is a no-op
  int requestor;
  int pd_enabled;
  int pd_spec_present;
  int eval_pods;          // pod lists present and evaluate_actuators
  int spec_cut_tile;      // speculation: tiles before this index assume every upgrade candidate gets a slot
  unsigned long long spec_sig;  // signature under which a device-resident hint from the previous call applies (0 = none)
  // sharding
  int rank;
  int world;
  // tiling
  int tile_nodes;         // nodes per tile (a multiple of 128, 128 .. UST_TILE_NODES)
  int n_tiles;            // tiles of the shard
  int tile_begin;         // streaming sub-range launches (pipelined uploads): tiles [tile_begin, tile_end)
  int tile_end;
  int static_rounds;      // rounds of the range a CTA takes in stride order before it claims tiles by ticket
  int parity;             // which accumulator set of the workspace this call uses (call number & 1)
  int seg;                // index of this streaming launch within the call (its ticket counter)
  int publish;            // split mode: this streaming launch is the last one of the call, its last CTA publishes P.xchg
  int split;              // split mode: a host-launched collective reduces P.xchg between the two kernels
  int relaxed;            // the call does not depend on the previous call of the handle (see apply_device): its streaming kernel
                          // starts without waiting for that call's verification kernel and overlaps its tail
  int stamps;             // diagnostics: write %globaltimer stamps
  // fused multi-GPU exchange (world > 1): mailboxes of all ranks as mapped into this process, call number
  int fused_exchange;
  long long epoch;
  UstMailbox* mbox[UST_MAX_WORLD];
};

// kernel launchers; all return cudaError_t as int. `pdl` = launch with programmatic stream serialization.
int ust_launch_stream(const UstParams& p, int grid, void* stream, int pdl);   // ust_stream.cu
int ust_launch_verify(const UstParams& p, int grid, void* stream, int pdl);   // ust_kernels.cu
int ust_stream_config(int device, int* num_sms, size_t* smem_bytes);          // also raises the dynamic shared-memory limit
int ust_launch_pod_summary(long long n, int active, const uint8_t* hot, const int32_t* pod_off, const uint16_t* pod_flags,
                           long long n_pods, const uint8_t* podlut, uint8_t* podsum, int grid, void* stream);
int ust_launch_build_state(long long n, const uint8_t* hot, const int32_t* ds_idx, int n_ds, const int32_t* ds_desired,
                           unsigned long long* ds_count, UstWorkspace* ws, ust_counters* out, int grid, void* stream);
// slot of a 128-bit UID in the DaemonSet hash table (before masking to the table size); host build and device lookup
#ifdef __CUDACC__
__host__ __device__
#endif
static inline unsigned ust_uid_hash(unsigned long long x, unsigned long long y) {
  unsigned h = (unsigned)x * 0x9E3779B9u + (unsigned)(x >> 32) * 0x85EBCA6Bu + (unsigned)y * 0xC2B2AE35u + (unsigned)(y >> 32) * 0x27D4EB2Fu;
  return h ^ (h >> 15);
}
int ust_launch_build_state_uids(long long n, const uint8_t* hot, const void* owner_uid, int n_ds, const void* ds_tab,
                                const int32_t* ds_tab_idx, int tab_slots, const int32_t* ds_desired, int32_t* ds_idx_out,
                                unsigned long long* ds_count, UstWorkspace* ws, ust_counters* out, int grid, void* stream);
int ust_launch_patch(long long m, const long long* idx, const uint8_t* state, const uint32_t* flags, const int32_t* pod_rev,
                     const int32_t* ds_idx, uint8_t* hot_out, uint32_t* flags_out, int32_t* rev_out, int32_t* ds_out, void* stream);
// rollout simulation: the clock of the feedback between two reconciles (include/ust.h, ust_sim_options)
struct UstSimParams {
  int timed;            // 0: whatever a node waits for has happened by the next reconcile
  long long now, dt;    // time of the reconcile that was just evaluated, seconds to the next one
  long long wait_timeout, job_seconds, validation_seconds, validation_timeout, maintenance_seconds;
};
int ust_launch_feedback(long long n, uint8_t* hot, uint32_t* flags, int32_t* pod_rev, const int32_t* ds_idx, int n_ds,
                        const int32_t* ds_rev, const uint8_t* next, const uint16_t* actions, const uint8_t* outcome,
                        const ust_counters* step, const UstSimParams& sp, int32_t* entered, int32_t* wait_start,
                        int32_t* valid_start, int grid, void* stream);
int ust_launch_sim_init(long long n, const uint32_t* flags, int32_t* entered, int32_t* wait_start, int32_t* valid_start, int grid,
                        void* stream);
int ust_launch_widen(long long n, const uint16_t* rev16, const int8_t* ds8, int32_t* rev_out, int32_t* ds_out, int grid,
                     void* stream);
// sparse outputs of a delta call: nodes whose (next_state, actions) differ from the previous call's, in node order
int ust_launch_diff(long long n, const uint8_t* next, const uint16_t* actions, const uint8_t* prev_next, const uint16_t* prev_actions,
                    unsigned int* block_count, long long* n_out, long long cap, long long* out_idx, uint8_t* out_next,
                    uint16_t* out_actions, void* stream);
int ust_diff_blocks(long long n);
