// ust_kernels.cu — the ApplyState kernels for sm_100a (B200).
//
// What replaces what: one launch of ust_fused_kernel computes, for every node of the snapshot, what the
// reference's ClusterUpgradeStateManagerImpl.ApplyState (pkg/upgrade/upgrade_state.go:171-281) computes
// with its twelve sequential Process* loops: next state label and actuator-call bitmask per node, plus
// the cluster counters of common_manager.go:715-788.
//
// Shape of the kernel (HBM-bound byte/integer streaming, no tensor-core work):
//   * cooperative grid of (#SM x 2 resident CTAs); CTA c owns one contiguous chunk of nodes, so slice order of the
//     upgrade-required bucket (upgrade_inplace.go:71) is chunk order;
//   * the per-policy transition table (32 KiB, built by the host: ust_lut.h) is staged into shared memory with one
//     TMA bulk copy that completes on an mbarrier; the first tile's loads are in flight before anything waits on it;
//   * ONE streaming pass over state(1 B) + flags(4) + pod_rev(4) + ds_idx(4) with 128-bit coalesced loads, four
//     steps (16 nodes) per thread in flight: every node is evaluated by one shared-memory lookup indexed by its hot
//     byte and one lookup in the transition table, and counted (14-bin state histogram, unavailable, upgrade
//     candidates) in byte-sliced SIMD-in-register counters; next_state(1) + actions(2) leave with full-width
//     coalesced stores: 16 algorithmic bytes per node, each touched once. The upgrade-slot grant - the only
//     cluster-wide dependency - is SPECULATED per chunk (from the policy, or from where the previous call's budget cut);
//   * one grid-wide barrier (single global atomic counter); every CTA then derives the slot budget
//     (GetUpgradesAvailable, common_manager.go:748-776) and checks the speculation in O(1); only when it cannot hold
//     do the CTAs scan the per-chunk candidate counts and re-evaluate the wrong interval, in pieces spread over the
//     whole grid (ordered path: hot-byte pre-pass + warp-scan ranks; uniform path elsewhere);
//   * last CTA out writes the counters and restores the workspace.
// Around it: ust_phase1/2_kernel (the same device code split at the barrier, for the NCCL mode and the pipelined
// host path), ust_pod_summary_kernel (pod lists -> one byte per node), ust_build_state*_kernel (BuildState),
// ust_patch_kernel / ust_feedback_kernel (delta updates, rollout simulation).
#include <cuda_runtime.h>

#include "ust_dev.h"

namespace {

constexpr int kThreads = UST_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kStep = kThreads * 4;   // nodes per CTA step (4 per thread)
#ifndef UST_UNROLL
#define UST_UNROLL 4
#endif
#ifndef UST_MIN_CTAS
#define UST_MIN_CTAS 2
#endif
constexpr int kUnroll = UST_UNROLL;            // steps in flight per thread in the fast path
constexpr int kTile = kStep * kUnroll;
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kScanSlots = 512;       // prefix table of the per-chunk candidate counts: one entry per chunk up to 512 chunks
constexpr int kMaxExactSteps = 256;   // steps per block of the exact (ordered) path: 256 x 1024 nodes
constexpr uint32_t kLutBytes = (UST_LUT_ENTRIES + 32) * sizeof(uint32_t);  // table + 16 {shift, base} pairs

struct __align__(128) Shared {
  uint32_t lut[UST_LUT_ENTRIES];  // + meta directly behind it: filled by ONE bulk (TMA) copy
  uint2 meta[16];
  uint4 hotent[256];             // per hot byte: {window shift - 2, table base, sixteen 4-bit one-hot count increments}
  int dsrev[UST_DS_SMEM_MAX + 1];
  unsigned int cnt[16];
  unsigned long long errinv;
  unsigned long long mbar;       // mbarrier the bulk copy completes on
  long long V[UST_V_LEN];
  // derived, CTA-uniform
  unsigned long long abort_key;  // ~0 = none
  long long budget;              // max(upgradesAvailable, 0)
  long long avail;
  long long max_unav;
  long long node_offset;         // global index of this shard's node 0
  long long cand_prefix;         // candidates before this chunk (global order)
  long long part[kWarps];
  int spec_cut;                  // effective speculative cut of this call (hint or default)
  int cut;                       // where the budget really cut: first chunk that is not fully granted (INT_MAX = none)
  int wrong_lo, wrong_hi;        // chunks whose speculation did not hold all lie in [wrong_lo, wrong_hi]
  long long tbase[kScanSlots];   // candidates before chunk slot * ceil(chunks / kScanSlots), shard-local
  unsigned int step_base[kMaxExactSteps];     // exact path: candidates before each step of the block
  unsigned char wtot[kMaxExactSteps][kWarps];  // ... and per warp within the step
  unsigned int chunk_cand;       // candidates of the chunk being streamed
  int next_chunk;                // next claimed chunk
  unsigned int warp_tot[kWarps];
  int last;
};

__device__ __forceinline__ uint4 ld_stream_u4(const void* p) { return __ldcs(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint32_t ld_keep_u32(const void* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// (pass + 1) of each state code, one nibble per code: position of its Process* pass in ApplyState
// (upgrade_state.go:205-274), 0 = never processed. Same content as ust_pass_of_state[] in ust_lut.h.
__device__ __forceinline__ int pass_of_state(unsigned code) {
  constexpr unsigned long long kPassPlus1 =
      (1ull << 0) | (3ull << 4) | (4ull << 8) | (5ull << 12) | (6ull << 16) | (7ull << 20) | (8ull << 24) | (0ull << 28) |
      (9ull << 32) | (11ull << 36) | (12ull << 40) | (2ull << 44) | (10ull << 48);
  return (int)((kPassPlus1 >> (4 * code)) & 15ull) - 1;
}

__device__ __forceinline__ long long chunk_bound(long long n, int c, int chunks) {
  if (c >= chunks) return n;
  long long b = (n * (long long)c) / chunks;
  return b & ~127LL;  // chunks start on 128-node boundaries: a warp's 128-node span never straddles two chunks
}

// ------------------------------------------------------------------------------------------------
// table staging: the per-policy transition table (DriverUpgradePolicySpec + manager options, compiled
// to 32 KiB by ust_lut.h) goes global -> shared with one TMA bulk copy that completes on an mbarrier;
// nothing waits for it until the first tile's loads are in flight.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stamp(const UstParams& P, int k) {
  if (threadIdx.x == 0 && k < 8) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.ws->dbg[blockIdx.x][k] = t;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ void p1_build_table(Shared& S);

__device__ void stage_tables_begin(const UstParams& P, Shared& S) {
  const int t = threadIdx.x;
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&S.mbar)), "r"(kLutBytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(S.lut)), "l"(P.lut), "r"(kLutBytes), "r"(smem_u32(&S.mbar)) : "memory");
  }
  if (P.n_ds <= UST_DS_SMEM_MAX)
    for (int i = t; i <= P.n_ds; i += kThreads) S.dsrev[i] = i < P.n_ds ? __ldg(P.ds_rev + i) : 0;
  p1_build_table(S);
  if (t < 16) S.cnt[t] = 0;
  if (t == 32) {
    S.errinv = 0; S.abort_key = ~0ull; S.chunk_cand = 0;
    // speculative cut: the previous call's, when it was made under the same signature; else the policy default
    const bool hinted = P.spec_sig != 0 && __ldcg(&P.ws->hint_sig) == P.spec_sig;
    S.spec_cut = hinted ? __ldcg(&P.ws->hint_cut) : P.spec_cut_chunk;
  }
}

__device__ __forceinline__ void stage_tables_wait(Shared& S) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(&S.mbar)) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// counting (part of the streaming pass)
//
// Byte-sliced SIMD-in-register counting. A 256-entry shared-memory table maps a hot byte to sixteen
// 4-bit one-hot increments packed in 64 bits (fields 0-13: state code, 14: unavailable, 15: upgrade
// candidate) next to the node's table window; a thread sums the entries of 8 nodes (no field can exceed 8),
// widens the nibbles to byte lanes, and keeps going. No atomics until the end of the chunk loop: one REDUX
// per counter per warp, 16 global atomics per CTA.
// ------------------------------------------------------------------------------------------------
// window shift (minus 2) of every state code, 8 bits each — compile-time copy of ust_window_shift[]
constexpr unsigned long long pack_shifts(int from) {
  unsigned long long v = 0;
  for (int i = 0; i < 8; i++) v |= (unsigned long long)(ust_window_shift[from + i] - 2) << (8 * i);
  return v;
}
constexpr unsigned long long kShiftLo = pack_shifts(0), kShiftHi = pack_shifts(8);

__device__ void p1_build_table(Shared& S) {
  // GetCurrentUnavailableNodes (common_manager.go:146-165) counts every snapshot entry that is cordoned or
  // not ready; an upgrade candidate is upgrade-required and not marked skip (upgrade_inplace.go:82)
  const unsigned b = threadIdx.x, code = b & 15u;
  unsigned long long v = 0;
  if (code < 14) {
    v = 1ull << (4 * code);
    if (b & (UST_HOT_UNSCHEDULABLE | UST_HOT_NOT_READY)) v |= 1ull << 56;
    if (code == UST_STATE_UPGRADE_REQUIRED && !(b & UST_HOT_SKIP)) v |= 1ull << 60;
  }
  const unsigned shift = (unsigned)(((code < 8 ? kShiftLo : kShiftHi) >> (8 * (code & 7))) & 0xFFull);
  S.hotent[b] = make_uint4(shift, code * (UST_LUT_WINDOW * 4u), (uint32_t)v, (uint32_t)(v >> 32));
}

// byte lanes: B[0] = fields 0,2,4,6  B[1] = fields 1,3,5,7  B[2] = fields 8,10,12,14  B[3] = fields 9,11,13,15
__device__ __forceinline__ unsigned p1_field(const uint32_t (&B)[4], int f) {
  return (B[(f >> 3) * 2 + (f & 1)] >> (8 * ((f & 7) >> 1))) & 0xFFu;
}

// ------------------------------------------------------------------------------------------------
// between the phases: cluster-wide scalars from the exchange vector (every CTA, redundantly)
// ------------------------------------------------------------------------------------------------
__device__ void derive_scalars(const UstParams& P, Shared& S) {
  const long long* V = S.V;
  const long long h0 = V[0], h1 = V[1], h2 = V[2], h4 = V[4], h11 = V[11];
  // GetTotalManagedNodes (common_manager.go:715-730): 11 buckets — not 6, 7, other
  const long long total = h0 + h1 + h2 + V[3] + h4 + V[5] + V[8] + V[9] + V[10] + h11 + V[12];
  const long long in_progress = total - h0 - h11 - h1;  // GetUpgradesInProgress (:733-739)
  unsigned long long abort_key = ~0ull;
  long long off = 0, my_off = 0, cand_before = 0;
  for (int r = 0; r < P.world; r++) {
    if (r == P.rank) my_off = off;
    if (r < P.rank) cand_before += V[UST_V_RANK_CAND + r];
    const unsigned long long e = (unsigned long long)V[UST_V_RANK_ERRINV + r];
    if (e) {
      const unsigned long long k = ~e;
      const unsigned long long gk = (k & 0xFF00000000000000ull) | ((k & 0x00FFFFFFFFFFFFFFull) + (unsigned long long)off);
      if (gk < abort_key) abort_key = gk;
    }
    off += V[UST_V_RANK_NODES + r];
  }
  long long max_unav = 0, avail = 0;
  const bool slots = P.active && !P.requestor;
  if (slots) {
    // upgrade_inplace.go:49-62 + intstr.GetScaledValueFromIntOrPercent(v, total, roundUp=true)
    if (P.max_unav_kind == UST_MAXUNAVAIL_INVALID && UST_KEY(2, 0) < abort_key) abort_key = UST_KEY(2, 0);
    max_unav = total;
    if (P.max_unav_kind == UST_MAXUNAVAIL_INT) max_unav = P.max_unav_value;
    else if (P.max_unav_kind == UST_MAXUNAVAIL_PERCENT)
      max_unav = (long long)ceil(__ddiv_rn(__dmul_rn((double)P.max_unav_value, (double)total), 100.0));
    // GetUpgradesAvailable (common_manager.go:748-776)
    avail = (P.max_parallel == 0) ? h1 : P.max_parallel - in_progress;
    const long long cur_unav = V[UST_V_UNAVAILABLE] + h2;
    if (avail > max_unav) avail = max_unav;
    if (cur_unav >= max_unav) avail = 0;
    else if (max_unav < total && cur_unav + avail > max_unav) avail = max_unav - cur_unav;
  }
  // SchedulePodEviction with a nil DeletionSpec (pod_manager.go:125-134)
  if (P.active && P.pd_enabled && !P.pd_spec_present && h4 > 0 && UST_KEY(5, 0) < abort_key) abort_key = UST_KEY(5, 0);
  S.abort_key = abort_key;
  S.avail = avail;
  S.max_unav = max_unav;
  S.budget = avail > 0 ? avail : 0;
  S.node_offset = my_off;
  S.cand_prefix = cand_before;  // completed with the chunk prefix by the caller
}

__device__ void write_counters(const UstParams& P, const Shared& S) {
  ust_counters c;
  const long long* V = S.V;
  for (int i = 0; i < 16; i++) c.hist[i] = V[i];
  c.unavailable = V[UST_V_UNAVAILABLE];
  c.candidates = V[UST_V_CANDIDATES];
  c.total_managed = V[0] + V[1] + V[2] + V[3] + V[4] + V[5] + V[8] + V[9] + V[10] + V[11] + V[12];
  c.in_progress = c.total_managed - V[0] - V[11] - V[1];
  c.error_code = UST_OK;
  c.error_index = -1;
  c.error_pass = -1;
  if (S.abort_key != ~0ull) {
    const int pass = (int)(S.abort_key >> 56);
    const long long idx1 = (long long)(S.abort_key & 0x00FFFFFFFFFFFFFFull);
    c.error_pass = pass;
    c.error_index = idx1 - 1;
    c.error_code = idx1 ? UST_ERR_REVISION_HASH : (pass == 2 ? UST_ERR_MAX_UNAVAILABLE : UST_ERR_POD_DELETION_SPEC);
  }
  const bool slots = P.active && !P.requestor && !(c.error_code && c.error_pass < 2) && c.error_code != UST_ERR_MAX_UNAVAILABLE;
  c.max_unavailable = slots ? S.max_unav : 0;
  c.upgrades_available = slots ? S.avail : 0;
  for (int i = 0; i < 7; i++) c.reserved[i] = 0;
  if (__ldcg(&P.ws->comm_timeout)) { c.error_code = UST_ERR_COMM; c.error_index = -1; c.error_pass = -1; }
  c.reserved[0] = (long long)__ldcg(&P.ws->fixups);  // chunks the verification phase had to redo (diagnostic)
  *P.out = c;
}

// ------------------------------------------------------------------------------------------------
// per-node transition
// ------------------------------------------------------------------------------------------------
template <bool DS_SMEM>
__device__ __forceinline__ bool pod_synced(const UstParams& P, const Shared& S, int rev, uint32_t di) {
  // podRevisionHash == daemonsetRevisionHash (common_manager.go:318); a missing DaemonSet never matches
  if (DS_SMEM) return (di < (uint32_t)P.n_ds) && (rev == S.dsrev[min(di, (uint32_t)P.n_ds)]);
  return di < (uint32_t)P.n_ds && rev == __ldg(P.ds_rev + di);
}

// table entry for one node. hb = hot byte, extra = derived bits (slot grant, pod-list summaries)
template <bool DS_SMEM>
__device__ __forceinline__ uint32_t node_entry(const UstParams& P, const Shared& S, uint32_t hb, uint32_t fl, int rev,
                                               uint32_t di, uint32_t extra) {
  uint32_t w = (fl & UST_F_INPUT_MASK) | ((hb >> 3) & (UST_W_SKIP | UST_W_UNSCHEDULABLE)) | extra;
  if (pod_synced<DS_SMEM>(P, S, rev, di)) w |= UST_W_SYNCED;
  const uint2 m = S.meta[hb & 15u];
  const uint32_t off = (__funnelshift_r(w, 0u, m.x) & 0x7FCu) | m.y;
  return *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(S.lut) + off);
}

__device__ __forceinline__ uint32_t noop_entry(uint32_t hb) { return ((hb & 15u) << 16) | 0xFF000000u; }

// abort semantics: nodes the sequential passes had not reached when the reference returned its error
// stay untouched; the aborting node carries UST_A_ERROR; an abort inside ProcessPodRestartNodes also
// drops the restarts collected so far, SchedulePodsRestart is only called after the loop
// (common_manager.go:462-523).
__device__ __forceinline__ uint32_t apply_abort(const Shared& S, uint32_t ent, uint32_t hb, long long gidx) {
  const int pass = pass_of_state(hb & 15u);
  if (pass < 0) return ent;
  const unsigned long long key = UST_KEY(pass, (unsigned long long)gidx + 1ull);
  if (key >= S.abort_key) {
    ent = noop_entry(hb);
    if (key == S.abort_key) ent |= UST_A_ERROR;
  } else if (pass == 8 && (S.abort_key >> 56) == 8) {
    ent &= ~(uint32_t)UST_A_RESTART_DRIVER_POD;
  }
  return ent;
}

__device__ __forceinline__ void pack4(const uint32_t e[4], uint32_t& next4, uint2& act4, uint32_t& out4) {
  act4.x = __byte_perm(e[0], e[1], 0x5410);
  act4.y = __byte_perm(e[2], e[3], 0x5410);
  const uint32_t hi01 = __byte_perm(e[0], e[1], 0x7632);  // [e0.b2 e0.b3 e1.b2 e1.b3]
  const uint32_t hi23 = __byte_perm(e[2], e[3], 0x7632);
  next4 = __byte_perm(hi01, hi23, 0x6420);
  out4 = __byte_perm(hi01, hi23, 0x7531);
}

// One tile = kUnroll steps of kStep nodes; thread t owns nodes base + j*kStep + 4t .. +3. `lim` is a
// multiple of 128 (or the tile is full), so validity is uniform per warp and per j.
template <int U>
struct TileT {
  uint32_t h[U];
  uint32_t ps[U];  // pod-list summaries (one byte per node), PODS variants only
  uint4 f[U], r[U], d[U];
};
// The streaming fast path keeps kUnroll steps in flight per thread; the redo variants after the barrier keep
// kColdUnroll: with a four-step tile the compiler parks their tile on the stack (STL/LDL in the tile loop, seen with
// nvdisasm), with a smaller one most of it stays in registers. Measured (profiles/README.md): 1 / 2 / 4 steps ->
// hinted redo 49.0 / 49.7 / 51.5 us, unhinted 81 / 77 / 78 us.
#ifndef UST_COLD_UNROLL
#define UST_COLD_UNROLL 2
#endif
constexpr int kColdUnroll = UST_COLD_UNROLL;

// Chunk-relative addressing: the chunk's base pointers are CTA-uniform; a thread addresses 4-node groups
// with a 32-bit group index q (thread t of the CTA owns groups done/4 + j*kStepQ + t of a tile).
struct Cursor {
  const uint32_t* h;
  const uint32_t* ps;  // null when the call has no pod lists
  const uint4* f;
  const uint4* r;
  const uint4* d;
  uint32_t* nx;
  uint2* ac;
  uint32_t* oc;
  int q;
};
constexpr int kStepQ = kStep / 4;  // a step in units of 4-node groups
constexpr int kTileQ = kTile / 4;

__device__ __forceinline__ Cursor cursor_at(const UstParams& P, long long base) {
  Cursor c;
  c.h = reinterpret_cast<const uint32_t*>(P.hot + base);
  c.ps = P.podsum ? reinterpret_cast<const uint32_t*>(P.podsum + base) : nullptr;
  c.f = reinterpret_cast<const uint4*>(P.flags + base);
  c.r = reinterpret_cast<const uint4*>(P.pod_rev + base);
  c.d = reinterpret_cast<const uint4*>(P.ds_idx + base);
  c.nx = reinterpret_cast<uint32_t*>(P.next + base);
  c.ac = reinterpret_cast<uint2*>(P.actions + base);
  c.oc = P.outcome ? reinterpret_cast<uint32_t*>(P.outcome + base) : nullptr;
  c.q = threadIdx.x;
  return c;
}
template <int U>
__device__ __forceinline__ void cursor_advance(Cursor& c) { c.q += U * kStepQ; }

// `room` = nodes left in the chunk from this thread's first node of the tile; chunk ends are multiples of
// 128 nodes, so for a partial tile validity is uniform per warp and per step.
// PODS: 0 = the call has no pod lists, 1 = it has, 2 = decided at run time (out-of-line variants)
template <bool FULL, int PODS, int U>
__device__ __forceinline__ void tile_load(const Cursor& c, int room, TileT<U>& T) {
#pragma unroll
  for (int j = 0; j < U; j++) {
    // every element is assigned on every path, so that the tile stays in registers (no stack copy)
    const bool valid = FULL || j * kStep + 4 <= room;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    T.h[j] = valid ? __ldg(c.h + c.q + j * kStepQ) : 0x0E0E0E0Eu;
    T.ps[j] = (valid && (PODS == 1 || (PODS == 2 && c.ps != nullptr))) ? __ldcs(c.ps + c.q + j * kStepQ) : 0u;
    T.f[j] = valid ? __ldcs(c.f + c.q + j * kStepQ) : zero;
    T.r[j] = valid ? __ldcs(c.r + c.q + j * kStepQ) : zero;
    T.d[j] = valid ? __ldcs(c.d + c.q + j * kStepQ) : zero;
  }
}

// revision-hash error seen while streaming (flags word already in registers)
__device__ __forceinline__ void spec_error_byte(const UstParams& P, Shared& S, unsigned b, uint32_t fl, long long i) {
  const unsigned code = b & 15u;
  if (!(b & UST_HOT_REVISION_HASH_ERROR) || !P.active) return;
  if (!(code == UST_STATE_UNKNOWN || code == UST_STATE_DONE || code == UST_STATE_POD_RESTART_REQUIRED || code == UST_STATE_FAILED)) return;
  if (fl & UST_F_POD_ORPHANED) return;
  atomicMax(&S.errinv, ~UST_KEY(pass_of_state(code), (unsigned long long)i + 1ull));
}

__device__ __forceinline__ void widen(uint32_t& lo, uint32_t& hi, uint32_t (&B)[4]) {
  B[0] += lo & 0x0F0F0F0Fu;
  B[1] += (lo >> 4) & 0x0F0F0F0Fu;
  B[2] += hi & 0x0F0F0F0Fu;
  B[3] += (hi >> 4) & 0x0F0F0F0Fu;
  lo = hi = 0;
}

// per-thread counting state of the streaming phase (lives across the chunks a CTA claims)
struct Acc {
  uint32_t B[4];           // sixteen byte-lane counters
  int tiles;               // tiles accumulated since the last spill (byte lanes hold 255)
  unsigned cand_spilled;   // candidates already spilled: cand_spilled + field 15 of B is monotonic per thread
};

__device__ __forceinline__ void spill_thread(Shared& S, Acc& A) {
#pragma unroll
  for (int f = 0; f < 16; f++) {
    const unsigned v = p1_field(A.B, f);
    if (v) atomicAdd(&S.cnt[f], v);
  }
  A.cand_spilled += p1_field(A.B, 15);
  A.B[0] = A.B[1] = A.B[2] = A.B[3] = 0;
  A.tiles = 0;
}

// One node of the streaming pass. `xs` = the node's hot byte moved to bits 4..11 of a word (so it indexes
// the 16-byte hotent table directly), `ws` = its SKIP / UNSCHEDULABLE bits already at w positions 2, 3.
// pod-list summary byte -> the w bits it stands for (ust_pod_summary_kernel): bit 0 = a wait-selector pod is
// running, bits 1..3 = UST_W_PD_HAS / UST_W_PD_MISMATCH / UST_W_DRAIN_ERROR, bit 4 = the list overrides the
// pre-evaluated UST_F_WAIT_PODS_RUNNING of the flags word
__device__ __forceinline__ uint32_t pods_apply(uint32_t fl, uint32_t ps) {
  return (fl & ~((ps & 0x10u) << 12)) | ((ps & 1u) << 16) | ((ps & 0xEu) << 21);
}
static_assert(UST_F_WAIT_PODS_RUNNING == (1u << 16) && UST_W_PD_HAS == (UST_PODSUM_TO_DELETE << 21) &&
              UST_W_PD_MISMATCH == (UST_PODSUM_CANNOT_DELETE << 21) && UST_W_DRAIN_ERROR == (UST_PODSUM_DRAIN_ERROR << 21) &&
              UST_PODSUM_WAIT_RUNNING == 1u, "pod summary byte layout");

template <bool DS_SMEM>
__device__ __forceinline__ uint32_t stream_node(const UstParams& P, const Shared& S, uint32_t tab_off, uint32_t wbits,
                                                uint32_t fl, int rev, uint32_t di, uint32_t grant, uint32_t& lo,
                                                uint32_t& hi) {
  const uint4 m = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(S.hotent) + tab_off);
  lo += m.z;
  hi += m.w;
  uint32_t w = fl | wbits | grant;  // fl: input bits of the flags word (+ pod-list bits), masked by the caller
  if (pod_synced<DS_SMEM>(P, S, rev, di)) w |= UST_W_SYNCED;
  const uint32_t off = (__funnelshift_r(w, 0u, m.x) & 0x7FCu) | m.y;
  return *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(S.lut) + off);
}

// candidate bytes of a hot word: bit 7 of byte k set iff node k is upgrade-required and not marked skip
__device__ __forceinline__ uint32_t cand_mask4(uint32_t x) {
  const uint32_t y = (x & 0x2F2F2F2Fu) ^ 0x01010101u;  // zero byte <=> code == 1 && !SKIP
  return ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) & 0x80808080u;
}

struct ExactCtx {
  int s0;              // step index (within the block) of the tile's first step
  unsigned int limit;  // candidates of the block that get a slot (rank < limit), block-relative
};

// One tile of the streaming pass: evaluate every node (one table lookup) and write next_state / actions;
// COUNT adds the byte-sliced counting, EXACT replaces the chunk-uniform slot grant by the ordered one: a
// candidate's rank in slice order = candidates before its step (S.step_base) + before its warp within the step
// (S.wtot) + a warp-shuffle exclusive scan — no block barrier in the loop (upgrade_inplace.go:71-109).
template <bool FULL, bool DS_SMEM, bool OUTCOME, bool COUNT, bool EXACT, int PODS, int U>
__device__ __forceinline__ void spec_tile(const UstParams& P, Shared& S, const Cursor& c, int room, long long i0,
                                          const TileT<U>& T, uint32_t grant, uint32_t (&B)[4], ExactCtx ex) {
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int j = 0; j < U; j++) {
    if (FULL || j * kStep + 4 <= room) {
      const uint32_t x = T.h[j];
      if (COUNT && (x & 0x80808080u)) {  // rare
        const long long i = i0 + j * kStep;
        spec_error_byte(P, S, x & 0xFFu, T.f[j].x, i);
        spec_error_byte(P, S, (x >> 8) & 0xFFu, T.f[j].y, i + 1);
        spec_error_byte(P, S, (x >> 16) & 0xFFu, T.f[j].z, i + 2);
        spec_error_byte(P, S, x >> 24, T.f[j].w, i + 3);
      }
      uint32_t g[4] = {grant, grant, grant, grant};
      if (EXACT) {
        const uint32_t cm = cand_mask4(x);
        const unsigned tc = __popc(cm);
        unsigned incl = tc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned v = __shfl_up_sync(kFull, incl, o);
          if ((threadIdx.x & 31) >= o) incl += v;
        }
        const int s = ex.s0 + j, warp = threadIdx.x >> 5;
        unsigned rank = S.step_base[s] + incl - tc;
#pragma unroll
        for (int w = 0; w < kWarps; w++)
          if (w < warp) rank += S.wtot[s][w];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const unsigned isc = (cm >> (8 * k + 7)) & 1u;
          g[k] = (isc && rank < ex.limit) ? UST_W_GRANTED : 0u;
          rank += isc;
        }
      }
      constexpr uint32_t kW = UST_W_SKIP | UST_W_UNSCHEDULABLE;
      uint32_t e[4];
      uint32_t dlo = 0, dhi = 0;
      // without pod lists the derived pod bits of w are never set: mask them out of the flags word
      constexpr uint32_t kIn = UST_F_INPUT_MASK;
      uint32_t fl[4] = {T.f[j].x & kIn, T.f[j].y & kIn, T.f[j].z & kIn, T.f[j].w & kIn};
      if (PODS) {
        const uint32_t ps = T.ps[j];
        fl[0] = pods_apply(fl[0], ps & 0xFFu);
        fl[1] = pods_apply(fl[1], (ps >> 8) & 0xFFu);
        fl[2] = pods_apply(fl[2], (ps >> 16) & 0xFFu);
        fl[3] = pods_apply(fl[3], ps >> 24);
      }
      e[0] = stream_node<DS_SMEM>(P, S, (x << 4) & 0xFF0u, (x >> 3) & kW, fl[0], (int)T.r[j].x, T.d[j].x, g[0], dlo, dhi);
      e[1] = stream_node<DS_SMEM>(P, S, (x >> 4) & 0xFF0u, (x >> 11) & kW, fl[1], (int)T.r[j].y, T.d[j].y, g[1], dlo, dhi);
      e[2] = stream_node<DS_SMEM>(P, S, (x >> 12) & 0xFF0u, (x >> 19) & kW, fl[2], (int)T.r[j].z, T.d[j].z, g[2], dlo, dhi);
      e[3] = stream_node<DS_SMEM>(P, S, (x >> 20) & 0xFF0u, (x >> 27) & kW, fl[3], (int)T.r[j].w, T.d[j].w, g[3], dlo, dhi);
      if (COUNT) { lo += dlo; hi += dhi; }
      uint32_t next4, out4;
      uint2 act4;
      pack4(e, next4, act4, out4);
      __stcs(c.nx + c.q + j * kStepQ, next4);
      __stcs(c.ac + c.q + j * kStepQ, act4);
      if (OUTCOME) __stcs(c.oc + c.q + j * kStepQ, out4);
    }
    if (COUNT && (j & 1)) widen(lo, hi, B);  // at most 8 per nibble so far
  }
  if (COUNT && (U & 1)) widen(lo, hi, B);
}

// Pre-pass of the ordered path over one block of steps [blk0, blk1): hot bytes only (they are L2-resident or
// about to be), all loads in flight at once; leaves per-step/per-warp candidate totals and the per-step exclusive
// prefix in shared memory. Returns the block's candidate total; `before` gets the candidates among the nodes
// [count_from, blk0) (the part of the chunk that precedes the block; same batch of loads, no extra round trip).
__device__ unsigned exact_prepass(const UstParams& P, Shared& S, long long blk0, long long blk1, long long count_from,
                                  long long& before) {
  const int t = threadIdx.x, warp = t >> 5;
  const int nsteps = (int)((blk1 - blk0 + kStep - 1) / kStep);
  unsigned pre = 0;
  for (long long a = count_from; a < blk0; a += 8LL * 16 * kThreads) {  // 16 nodes per load, 8 loads in flight
    uint4 x[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const long long i = a + (long long)u * 16 * kThreads + 16 * t;     // count_from, blk0: multiples of 128
      x[u] = i < blk0 ? __ldg(reinterpret_cast<const uint4*>(P.hot + i)) : make_uint4(0x0E0E0E0Eu, 0x0E0E0E0Eu, 0x0E0E0E0Eu, 0x0E0E0E0Eu);
    }
#pragma unroll
    for (int u = 0; u < 8; u++)
      pre += __popc(cand_mask4(x[u].x)) + __popc(cand_mask4(x[u].y)) + __popc(cand_mask4(x[u].z)) + __popc(cand_mask4(x[u].w));
  }
  for (int s0 = 0; s0 < nsteps; s0 += 8) {
    uint32_t x[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const long long i = blk0 + (long long)(s0 + u) * kStep + 4 * t;
      x[u] = (s0 + u < nsteps && i + 4 <= blk1) ? __ldg(reinterpret_cast<const uint32_t*>(P.hot + i)) : 0x0E0E0E0Eu;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const unsigned w = __reduce_add_sync(kFull, (unsigned)__popc(cand_mask4(x[u])));
      if ((t & 31) == 0 && s0 + u < nsteps) S.wtot[s0 + u][warp] = (unsigned char)w;  // <= 128
    }
  }
  pre = __reduce_add_sync(kFull, pre);
  if ((t & 31) == 0) S.part[warp] = pre;
  __syncthreads();
  // exclusive prefix over the steps of the block (kMaxExactSteps == kThreads: one step per thread)
  unsigned tot = 0;
  if (t < nsteps)
    for (int w = 0; w < kWarps; w++) tot += S.wtot[t][w];
  unsigned incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(kFull, incl, o);
    if ((t & 31) >= o) incl += v;
  }
  if ((t & 31) == 31) S.warp_tot[warp] = incl;
  __syncthreads();
  unsigned before_steps = 0, total = 0;
  long long pre_all = 0;
#pragma unroll
  for (int w = 0; w < kWarps; w++) {
    const unsigned v = S.warp_tot[w];
    if (w < warp) before_steps += v;
    total += v;
    pre_all += S.part[w];
  }
  if (t < nsteps) S.step_base[t] = before_steps + incl - tot;
  __syncthreads();
  before = pre_all;
  return total;
}

// Evaluate the block [blk0, blk1) of a chunk through the 4-deep tile pipeline. ORDERED: `limit` slots are left
// for the candidates from node `count_from` (<= blk0) on, in slice order (pre-pass + per-node ranks); otherwise
// the slot grant is uniform (`grant`). Returns the candidates in [count_from, blk1) (ORDERED only).
template <bool DS_SMEM, bool OUTCOME, bool COUNT, bool ORDERED, int PODS, int U>
__device__ __forceinline__ long long spec_block(const UstParams& P, Shared& S, long long blk0, long long blk1, uint32_t grant,
                                                long long limit, long long count_from, Acc& A, bool wait_for_table) {
  const long long span = blk1 - blk0;  // CTA-uniform, a multiple of 128 (or <= 0 for an empty chunk)
  // the streaming fast path has a specialised body for full tiles; the out-of-line variants keep one (predicated) body
  constexpr bool kFullVariant = COUNT && !ORDERED;
  constexpr int kTileU = U * kStep;  // nodes per tile
  const int t4 = 4 * threadIdx.x;
  auto room_at = [&](long long done) -> int {  // nodes from this thread's first node of the tile to the block end
    const long long r = span - done - t4;
    return r > (1LL << 30) ? (1 << 30) : (r < 0 ? 0 : (int)r);
  };
  Cursor c = cursor_at(P, blk0);
  long long i0 = blk0 + t4;
  TileT<U> T;
  auto load = [&](long long done) {  // loads of the tile starting `done` nodes into the block; c.q points at it
    if (kFullVariant && span - done >= kTileU) tile_load<true, PODS, U>(c, 0, T); else tile_load<false, PODS, U>(c, room_at(done), T);
  };
  if (span > 0) load(0);  // the first tile's loads go out before anything waits (table copy, pre-pass)
  if (wait_for_table) {
    stage_tables_wait(S);
    __syncthreads();
  }
  ExactCtx ex{0, 0};
  long long blk_cand = 0;  // candidates in [count_from, blk1)
  if (ORDERED && span > 0) {
    long long before = 0;
    blk_cand = exact_prepass(P, S, blk0, blk1, count_from, before);
    stamp(P, 6);
    limit -= before;  // `limit` slots were left at count_from
    blk_cand += before;
    ex.limit = limit <= 0 ? 0u : (limit > 0x7FFFFFFFLL ? 0x7FFFFFFFu : (unsigned)limit);
  }
#pragma unroll 1
  for (long long done = 0; done < span; done += kTileU) {
    if (done) load(done);
    const int room = room_at(done);
    if (ORDERED) ex.s0 = (int)(done / kStep);
    if (kFullVariant && span - done >= kTileU) spec_tile<true, DS_SMEM, OUTCOME, COUNT, ORDERED, PODS, U>(P, S, c, room, i0, T, grant, A.B, ex);
    else spec_tile<false, DS_SMEM, OUTCOME, COUNT, ORDERED, PODS, U>(P, S, c, room, i0, T, grant, A.B, ex);
    cursor_advance<U>(c);
    i0 += kTileU;
    if (COUNT && ++A.tiles >= 224 / (4 * U)) spill_thread(S, A);  // byte lanes: at most 4U per tile, 255 max
  }
  return blk_cand;
}

// The streaming fast path: uniform grant, counting, inlined into the chunk loop.
template <bool DS_SMEM, bool OUTCOME, bool PODS>
__device__ __forceinline__ void spec_chunk(const UstParams& P, Shared& S, long long b0, long long lim, uint32_t grant, Acc& A,
                                           bool wait_for_table) {
  spec_block<DS_SMEM, OUTCOME, true, false, PODS ? 1 : 0, kUnroll>(P, S, b0, lim, grant, 0, b0, A, wait_for_table);
}

// The ordered variant lives out of line so that it cannot cost the fast path registers: blocks of
// kMaxExactSteps steps; `slots` slots are left for the candidates from node `count_from` (<= b0, same chunk) on.
// Counts (COUNT) go straight to the CTA's shared counters; returns the candidates in [count_from, lim).
template <bool DS_SMEM, bool OUTCOME, bool COUNT>
__device__ __forceinline__ long long ordered_range(const UstParams& P, Shared& S, long long b0, long long lim, long long slots,
                                                long long count_from, bool wait_for_table) {
  Acc A;
  A.B[0] = A.B[1] = A.B[2] = A.B[3] = 0;
  A.tiles = 0;
  A.cand_spilled = 0;
  constexpr long long kBlk = (long long)kMaxExactSteps * kStep;
  long long seen = 0;
  long long blk0 = b0;
  do {
    const long long blk1 = blk0 + kBlk < lim ? blk0 + kBlk : lim;
    seen += spec_block<DS_SMEM, OUTCOME, COUNT, true, 2, kColdUnroll>(P, S, blk0, blk1, 0u, slots - seen, blk0 == b0 ? count_from : blk0, A,
                                                      wait_for_table);
    wait_for_table = false;
    __syncthreads();  // step_base / wtot are rewritten by the next block's pre-pass
    blk0 = blk1;
  } while (blk0 < lim);
  if (COUNT) spill_thread(S, A);
  return seen;
}

// ... and so does the uniform-grant variant without counting (redo of chunks whose every candidate, or none, gets a slot)
template <bool DS_SMEM, bool OUTCOME>
__device__ __forceinline__ void uniform_range(const UstParams& P, Shared& S, long long b0, long long lim, uint32_t grant) {
  Acc A;
  A.B[0] = A.B[1] = A.B[2] = A.B[3] = 0;
  A.tiles = 0;
  A.cand_spilled = 0;
  spec_block<DS_SMEM, OUTCOME, false, false, 2, kColdUnroll>(P, S, b0, lim, grant, 0, b0, A, false);
}

// general path: one step of kStep nodes, bounds-checked; optional exact ordered slot allocation,
// abort masking and pod-list evaluation
template <bool EXACT>
__device__ void general_step(const UstParams& P, Shared& S, long long base, long long b1, uint32_t grant,
                             long long& running /* candidates seen in this chunk so far (EXACT) */) {
  const int t = threadIdx.x;
  const long long i0 = base + 4 * t;
  const bool aborting = S.abort_key != ~0ull;
  uint32_t hb[4], fl[4], di[4];
  int rev[4];
  int nvalid = 0;
  if (i0 + 4 <= b1) {
    nvalid = 4;
    const uint32_t h = ld_keep_u32(P.hot + i0);
    const uint4 f = ld_stream_u4(P.flags + i0), r = ld_stream_u4(P.pod_rev + i0), d = ld_stream_u4(P.ds_idx + i0);
    hb[0] = h & 0xFFu; hb[1] = (h >> 8) & 0xFFu; hb[2] = (h >> 16) & 0xFFu; hb[3] = h >> 24;
    fl[0] = f.x; fl[1] = f.y; fl[2] = f.z; fl[3] = f.w;
    rev[0] = (int)r.x; rev[1] = (int)r.y; rev[2] = (int)r.z; rev[3] = (int)r.w;
    di[0] = d.x; di[1] = d.y; di[2] = d.z; di[3] = d.w;
  } else if (i0 < b1) {
    nvalid = (int)(b1 - i0);
    for (int k = 0; k < 4; k++) {
      const bool v = k < nvalid;
      hb[k] = v ? P.hot[i0 + k] : (uint32_t)UST_STATE_EXCLUDED;
      fl[k] = v ? P.flags[i0 + k] : 0u;
      rev[k] = v ? P.pod_rev[i0 + k] : 0;
      di[k] = v ? (uint32_t)P.ds_idx[i0 + k] : 0xFFFFFFFFu;
    }
  } else {
    for (int k = 0; k < 4; k++) { hb[k] = UST_STATE_EXCLUDED; fl[k] = 0; rev[k] = 0; di[k] = 0xFFFFFFFFu; }
  }

  uint32_t gbits[4] = {grant, grant, grant, grant};
  if (EXACT) {
    // ordered slot allocation: candidate = upgrade-required && !skip; rank = exclusive count of
    // candidates in slice order; granted iff rank < max(upgradesAvailable, 0) (upgrade_inplace.go:71-109)
    unsigned c[4], tc = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      c[k] = ((hb[k] & 15u) == UST_STATE_UPGRADE_REQUIRED && !(hb[k] & UST_HOT_SKIP)) ? 1u : 0u;
      tc += c[k];
    }
    unsigned incl = tc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(kFull, incl, o);
      if ((t & 31) >= o) incl += v;
    }
    if ((t & 31) == 31) S.warp_tot[t >> 5] = incl;
    __syncthreads();
    unsigned before = 0, step_total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      const unsigned v = S.warp_tot[w];
      if (w < (t >> 5)) before += v;
      step_total += v;
    }
    __syncthreads();
    long long rank = S.cand_prefix + running + before + (incl - tc);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      gbits[k] = (c[k] && rank < S.budget) ? UST_W_GRANTED : 0u;
      rank += c[k];
    }
    running += step_total;
  }

  if (nvalid == 0) return;
  const bool ds_smem = P.n_ds <= UST_DS_SMEM_MAX;
  uint32_t e[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint32_t extra = gbits[k], f = fl[k];
    if (P.podsum && k < nvalid) {  // pod-list summary of the node (ust_pod_summary_kernel), see pods_apply()
      const uint32_t ps = P.podsum[i0 + k];
      f &= ~((ps & 0x10u) << 12);
      extra |= ((ps & 1u) << 16) | ((ps & 0xEu) << 21);
    }
    e[k] = ds_smem ? node_entry<true>(P, S, hb[k], f, rev[k], di[k], extra) : node_entry<false>(P, S, hb[k], f, rev[k], di[k], extra);
    if (aborting) e[k] = apply_abort(S, e[k], hb[k], S.node_offset + i0 + k);
  }
  uint32_t next4, out4;
  uint2 act4;
  pack4(e, next4, act4, out4);
  if (nvalid == 4) {
    __stcs(reinterpret_cast<uint32_t*>(P.next + i0), next4);
    __stcs(reinterpret_cast<uint2*>(P.actions + i0), act4);
    if (P.outcome) __stcs(reinterpret_cast<uint32_t*>(P.outcome + i0), out4);
  } else {
    for (int k = 0; k < nvalid; k++) {
      P.next[i0 + k] = (uint8_t)(e[k] >> 16);
      P.actions[i0 + k] = (uint16_t)e[k];
      if (P.outcome) P.outcome[i0 + k] = (uint8_t)(e[k] >> 24);
    }
  }
}

// how many candidates of a chunk get a slot, given the cluster-wide budget and the candidates before the chunk
__device__ __forceinline__ long long required_local(const Shared& S, unsigned chunk_cand) {
  long long l = S.budget - S.cand_prefix;
  if (l < 0) l = 0;
  return l > (long long)chunk_cand ? (long long)chunk_cand : l;
}
__device__ __forceinline__ uint32_t spec_grant(const UstParams& P, const Shared& S, int chunk) {
  return (P.active && !P.requestor && chunk < S.spec_cut) ? UST_W_GRANTED : 0u;
}
// Redo one chunk through the bounds-checked step path: aborts (nodes of later passes keep their state) and
// pod lists (per-node CSR walk) need it; the slot grant is exact here too.
__device__ void general_chunk(const UstParams& P, Shared& S, long long b0, long long b1, unsigned chunk_cand) {
  const bool slotted = P.active && !P.requestor;
  const long long need = slotted ? required_local(S, chunk_cand) : 0;
  const bool ordered = slotted && chunk_cand != 0 && need > 0 && need < (long long)chunk_cand;
  const uint32_t grant = (slotted && chunk_cand != 0 && need == (long long)chunk_cand) ? UST_W_GRANTED : 0u;
  long long running = 0;
  if (!ordered) {
    for (long long base = b0; base < b1; base += kStep) general_step<false>(P, S, base, b1, grant, running);
  } else {
    for (long long base = b0; base < b1; base += kStep) general_step<true>(P, S, base, b1, 0u, running);
  }
}

// The chunk loop of the streaming phase: chunks (contiguous node ranges, chunk order == slice order) are
// claimed with an atomic ticket, so CTAs that HBM serves faster take more of them and all CTAs reach the
// grid barrier together. Per chunk: counts + speculative outputs; the chunk's candidate count is published
// for the ordered slot allocation. Returns the number of nodes this CTA streamed.
template <bool DS_SMEM, bool OUTCOME, bool PODS>
__device__ long long stream_loop(const UstParams& P, Shared& S) {
  const int t = threadIdx.x;
  const int n_chunks = P.grid_chunks;
  UstWorkspace* ws = P.ws;
  long long nodes_seen = 0;
  Acc A;
  A.B[0] = A.B[1] = A.B[2] = A.B[3] = 0;
  A.tiles = 0;
  A.cand_spilled = 0;
  bool first = true;
  unsigned cand_published = 0;
  int chunk = P.chunk_begin + blockIdx.x;
  while (chunk < P.chunk_end) {
    if (t == 0) S.next_chunk = P.chunk_begin + (int)(atomicAdd(&ws->ticket, 1u) + gridDim.x);  // claimed early: its latency hides behind the chunk
    const long long b0 = chunk_bound(P.n, chunk, n_chunks), b1 = chunk_bound(P.n, chunk + 1, n_chunks);
    const long long lim = b1 & ~127LL;  // == b1 except for the ragged end of the whole array
    const uint32_t grant = spec_grant(P, S, chunk);
    const unsigned cand0 = A.cand_spilled + p1_field(A.B, 15);
    spec_chunk<DS_SMEM, OUTCOME, PODS>(P, S, b0, lim, grant, A, first);
    first = false;
    if (lim < b1) {  // ragged end (< 128 nodes, last chunk only)
      long long running = 0;
      general_step<false>(P, S, lim, b1, grant, running);
      for (long long j = lim + t; j < b1; j += kThreads) {
        const unsigned b = P.hot[j];
        uint32_t lo = S.hotent[b].z, hi = S.hotent[b].w;
        widen(lo, hi, A.B);
        spec_error_byte(P, S, b, P.flags[j], j);
      }
    }
    const unsigned mine = A.cand_spilled + p1_field(A.B, 15) - cand0;
    const unsigned warp_cand = __reduce_add_sync(kFull, mine);
    if ((t & 31) == 0 && warp_cand) atomicAdd(&S.chunk_cand, warp_cand);
    __syncthreads();
    const int next = S.next_chunk;
    if (t == 0) {  // S.chunk_cand only ever grows: no shared write between the two barriers
      const unsigned total = S.chunk_cand;
      ws->cand_cta[chunk] = total - cand_published;
      cand_published = total;
    }
    __syncthreads();
    nodes_seen += b1 - b0;
    chunk = next;
  }
  if (first) stage_tables_wait(S);  // no chunk for this CTA: still never exit with the bulk copy in flight
  // all threads converged: one REDUX per counter per warp, one shared atomic per warp
#pragma unroll
  for (int f = 0; f < 16; f++) {
    const unsigned v = __reduce_add_sync(kFull, p1_field(A.B, f));
    if ((t & 31) == 0 && v) atomicAdd(&S.cnt[f], v);
  }
  return nodes_seen;
}

// Streaming phase of one CTA.
__device__ void stream_phase(const UstParams& P, Shared& S) {
  const int t = threadIdx.x;
  UstWorkspace* ws = P.ws;
  long long nodes_seen = 0;
  const int variant = (P.n_ds <= UST_DS_SMEM_MAX ? 4 : 0) | (P.outcome ? 2 : 0) | (P.podsum ? 1 : 0);
  switch (variant) {
    case 7: nodes_seen = stream_loop<true, true, true>(P, S); break;
    case 6: nodes_seen = stream_loop<true, true, false>(P, S); break;
    case 5: nodes_seen = stream_loop<true, false, true>(P, S); break;
    case 4: nodes_seen = stream_loop<true, false, false>(P, S); break;
    case 3: nodes_seen = stream_loop<false, true, true>(P, S); break;
    case 2: nodes_seen = stream_loop<false, true, false>(P, S); break;
    case 1: nodes_seen = stream_loop<false, false, true>(P, S); break;
    default: nodes_seen = stream_loop<false, false, false>(P, S); break;
  }
  __syncthreads();
  // 16 global atomics per CTA
  if (t < 14) {
    if (S.cnt[t]) atomicAdd(&ws->acc[t], (unsigned long long)S.cnt[t]);
  } else if (t == 14) {
    unsigned long long in = 0;
    for (int f = 0; f < 14; f++) in += S.cnt[f];
    const unsigned long long excluded = (unsigned long long)nodes_seen - in;
    if (excluded) atomicAdd(&ws->acc[UST_STATE_EXCLUDED], excluded);
  } else if (t == 15) {
    if (S.cnt[14]) atomicAdd(&ws->acc[UST_V_UNAVAILABLE], (unsigned long long)S.cnt[14]);
    if (S.cnt[15]) atomicAdd(&ws->acc[UST_V_CANDIDATES], (unsigned long long)S.cnt[15]);
  } else if (t == 32) {
    if (S.errinv) atomicMax(&ws->errinv, S.errinv);
  }
}

__device__ void load_local_vector(const UstParams& P, Shared& S) {
  // world == 1: the exchange vector is just this shard's accumulators
  const int t = threadIdx.x;
  if (t < UST_V_LEN) {
    long long v = 0;
    if (t < 18) v = (long long)__ldcg(&P.ws->acc[t]);
    else if (t == UST_V_RANK_CAND + P.rank) v = (long long)__ldcg(&P.ws->acc[UST_V_CANDIDATES]);
    else if (t == UST_V_RANK_NODES + P.rank) v = P.n;
    else if (t == UST_V_RANK_ERRINV + P.rank) v = (long long)__ldcg(&P.ws->errinv);
    S.V[t] = v;
  }
}

// ---- system-scope accessors for the NVLink mailbox exchange -------------------------------------------------
__device__ __forceinline__ void st_relaxed_sys(long long* p, long long v) { asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_release_sys(long long* p, long long v) { asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long ld_relaxed_sys(const long long* p) {
  long long v;
  asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kCommTimeoutNs = 4000000000ull;  // give up on a missing peer after 4 s

// Cluster-wide vector for world > 1 without leaving the kernel: CTA 0 waits for the local CTAs, pushes this
// shard's lanes into every rank's mailbox over NVLink, waits for every rank's lanes in its own mailbox, sums,
// and publishes the result to the other local CTAs. One-hot per-rank lanes make the sum an all-gather.
__device__ void fused_exchange(const UstParams& P, Shared& S) {
  const int t = threadIdx.x;
  UstWorkspace* ws = P.ws;
  const int par = (int)(P.epoch & 1);
  if (blockIdx.x == 0) {
    if (t == 0) {
      while (ld_acquire_u32(&ws->arrive) < gridDim.x) __nanosleep(20);
      __threadfence();
    }
    __syncthreads();
    load_local_vector(P, S);
    __syncthreads();
    if (t < UST_V_LEN) {
      for (int r = 0; r < P.world; r++) st_relaxed_sys(&P.mbox[r]->slot[par][P.rank][t], S.V[t]);
      __threadfence_system();
    }
    __syncthreads();
    if (t < P.world) st_release_sys(&P.mbox[t]->slot[par][P.rank][UST_MBOX_FLAG], P.epoch);
    if (t < P.world) {
      const unsigned long long t0 = now_ns();
      while (ld_acquire_sys(&P.mbox[P.rank]->slot[par][t][UST_MBOX_FLAG]) != P.epoch) {
        if (now_ns() - t0 > kCommTimeoutNs) { ws->comm_timeout = 1; break; }
        __nanosleep(50);
      }
    }
    __syncthreads();
    if (t < UST_V_LEN) {
      long long sum = 0;
      for (int r = 0; r < P.world; r++) sum += ld_relaxed_sys(&P.mbox[P.rank]->slot[par][r][t]);
      S.V[t] = sum;
      ws->gv[t] = sum;
    }
    __threadfence();
    __syncthreads();
    if (t == 0) {
      asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&ws->go), "l"((unsigned long long)P.epoch) : "memory");
    }
  } else {
    if (t == 0) {
      const unsigned long long t0 = now_ns();
      while (ld_acquire_u64(&ws->go) != (unsigned long long)P.epoch) {
        if (now_ns() - t0 > kCommTimeoutNs + 1000000000ull) break;
        __nanosleep(50);
      }
      __threadfence();
    }
    __syncthreads();
    if (t < UST_V_LEN) S.V[t] = __ldcg(&ws->gv[t]);
  }
  __syncthreads();
}

__device__ __forceinline__ bool verification_needed(const UstParams& P, const Shared& S);

__device__ void finish(const UstParams& P, Shared& S, bool reset_ws) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(&P.ws->depart, 1u);
    if (prev == gridDim.x - 1u) {  // last CTA out: publish counters, restore the workspace invariant
      write_counters(P, S);
      if (P.spec_sig != 0 && P.active && !P.requestor && S.abort_key == ~0ull) {
        // where the budget really cut this time = next call's speculation
        const int cut = verification_needed(P, S) ? S.cut : S.spec_cut;  // else: all-or-nothing guess that held
        P.ws->hint_cut = cut;
        P.ws->hint_sig = P.spec_sig;
      }
      P.ws->fixups = 0;
      P.ws->comm_timeout = 0;
      if (reset_ws) {
        for (int i = 0; i < 18; i++) P.ws->acc[i] = 0;
        P.ws->errinv = 0;
        P.ws->ticket = 0;
      }
      P.ws->arrive = 0;
      P.ws->depart = 0;
      __threadfence();
    }
  }
}

// Is any chunk's speculation possibly wrong? O(1) from the cluster-wide scalars: with "nobody gets a
// slot" the speculation only fails if there is a budget at all, with "everybody gets one" only if the
// budget is smaller than the number of candidates.
__device__ __forceinline__ bool verification_needed(const UstParams& P, const Shared& S) {
  if (S.abort_key != ~0ull) return true;
  if (!(P.active && !P.requestor)) return false;
  const long long cands = S.V[UST_V_CANDIDATES];
  if (cands == 0) return false;
  if (S.spec_cut > 0 && S.spec_cut < P.grid_chunks) return true;  // hint in the middle: check chunk by chunk
  return S.spec_cut ? (S.budget < cands) : (S.budget > 0);
}

// Per-chunk candidate counts of the chunks a thread scans: thread t owns the chunks [t * cpt, (t + 1) * cpt),
// cpt = 2 * ceil(chunks / kScanSlots). Loaded right after the grid barrier, together with the exchange vector.
constexpr int kScanRegs = 4;
struct ChunkCands { unsigned int v[kScanRegs]; };
__device__ __forceinline__ int scan_slot_chunks(const UstParams& P) { return (P.grid_chunks + kScanSlots - 1) / kScanSlots; }
__device__ __forceinline__ ChunkCands load_chunk_cands(const UstParams& P) {
  ChunkCands r;
  const int cpt = 2 * scan_slot_chunks(P), c0 = threadIdx.x * cpt;
#pragma unroll
  for (int k = 0; k < kScanRegs; k++) r.v[k] = (k < cpt && c0 + k < P.grid_chunks) ? __ldcg(&P.ws->cand_cta[c0 + k]) : 0u;
  return r;
}

// Every CTA scans the per-chunk candidate counts once: where the budget really cuts, and the interval of chunks
// whose speculation did not hold. Leaves prefix bases in shared memory for chunk_prefix().
__device__ void scan_chunks(const UstParams& P, Shared& S, long long rank_base, const ChunkCands& held) {
  const int t = threadIdx.x, n_chunks = P.grid_chunks;
  const int spt = scan_slot_chunks(P), cpt = 2 * spt;
  const int c0 = t * cpt < n_chunks ? t * cpt : n_chunks, c1 = c0 + cpt < n_chunks ? c0 + cpt : n_chunks;
  const bool in_regs = cpt <= kScanRegs;
  auto cand_of = [&](int c) -> long long { return __ldcg(&P.ws->cand_cta[c]); };
  if (t == 0) { S.wrong_lo = 0x7FFFFFFF; S.wrong_hi = -1; S.cut = 0x7FFFFFFF; }
  long long mine = 0;
  if (in_regs) {
#pragma unroll
    for (int k = 0; k < kScanRegs; k++) mine += held.v[k];
  } else {
    for (int c = c0; c < c1; c++) mine += cand_of(c);
  }
  long long incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long v = __shfl_up_sync(kFull, incl, o);
    if ((t & 31) >= o) incl += v;
  }
  if ((t & 31) == 31) S.part[t >> 5] = incl;
  __syncthreads();
  long long before = 0;
#pragma unroll
  for (int w = 0; w < kWarps; w++)
    if (w < (t >> 5)) before += S.part[w];
  const bool slotted = P.active && !P.requestor;
  int lo = 0x7FFFFFFF, hi = -1, cut = 0x7FFFFFFF;
  long long local = before + incl - mine;  // shard-local candidates before chunk c
  auto visit = [&](int c, long long cand) {
    if ((c - c0) % spt == 0) S.tbase[c / spt] = local;
    const long long pre = rank_base + local;
    if (cand != 0 && slotted) {
      long long req = S.budget - pre;
      req = req < 0 ? 0 : (req > cand ? cand : req);
      const long long spec = c < S.spec_cut ? cand : 0;
      if (req != spec) { lo = lo < c ? lo : c; hi = c; }
      if (pre + cand > S.budget && c < cut) cut = c;
    }
    local += cand;
  };
  if (in_regs) {
#pragma unroll
    for (int k = 0; k < kScanRegs; k++)
      if (c0 + k < c1) visit(c0 + k, held.v[k]);
  } else {
    for (int c = c0; c < c1; c++) visit(c, cand_of(c));
  }
  lo = __reduce_min_sync(kFull, lo);
  hi = __reduce_max_sync(kFull, hi);
  cut = __reduce_min_sync(kFull, cut);
  if ((t & 31) == 0) {
    if (hi >= 0) { atomicMin(&S.wrong_lo, lo); atomicMax(&S.wrong_hi, hi); }
    atomicMin(&S.cut, cut);
  }
  __syncthreads();
}

// shard-local candidates before chunk c (after scan_chunks); no loads up to kScanSlots chunks
__device__ __forceinline__ long long chunk_prefix(const UstParams& P, const Shared& S, int c) {
  const int spt = scan_slot_chunks(P);
  const int slot = c / spt;
  long long v = S.tbase[slot];
  for (int k = slot * spt; k < c; k++) v += __ldcg(&P.ws->cand_cta[k]);
  return v;
}

// pull the first tile of a piece towards L2 while its starting rank is still being worked out
__device__ __forceinline__ void prefetch_piece(const UstParams& P, long long p0, long long p1) {
  const long long end = p1 - p0 > kTile ? p0 + kTile : p1;
  for (long long i = p0 + 32LL * threadIdx.x; i < end; i += 32LL * kThreads) {  // 32 nodes = one 128-byte line of an int32 array
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.flags + i));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.pod_rev + i));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.ds_idx + i));
  }
}

// The slot speculation was wrong for the chunks in [wrong_lo, wrong_hi] (a contiguous node range): re-evaluate
// them with the ordered grant (global candidate rank < budget). The range is cut into pieces that are spread
// over ALL CTAs of the grid - a single CTA would be latency-bound - each piece finding its starting rank from
// the per-chunk counts plus a hot-byte count of the part of its chunk that precedes it.
__device__ void redo_wrong_chunks(const UstParams& P, Shared& S, long long rank_base) {
  const int lo = S.wrong_lo, hi = S.wrong_hi, n_chunks = P.grid_chunks;
  if (lo > hi) return;
  const int m = hi - lo + 1;
  const long long max_len = P.n / n_chunks + 256;                 // no chunk is longer (chunk_bound rounds to 128)
  const int max_steps = (int)((max_len + kStep - 1) / kStep);
  int pieces = (int)gridDim.x / m;                                  // at most one piece per CTA when that is possible
  pieces = pieces < 1 ? 1 : (pieces > max_steps ? max_steps : pieces);
  const int variant = (P.n_ds <= UST_DS_SMEM_MAX ? 2 : 0) | (P.outcome ? 1 : 0);
  const long long items = (long long)m * pieces;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int chunk = lo + (int)(item / pieces), j = (int)(item % pieces);
    const long long b0 = chunk_bound(P.n, chunk, n_chunks), b1 = chunk_bound(P.n, chunk + 1, n_chunks);
    const long long lim = b1 & ~127LL;
    const long long steps = (b1 - b0 + kStep - 1) / kStep, per = (steps + pieces - 1) / pieces;
    const long long p0 = b0 + (long long)j * per * kStep;
    long long p1 = p0 + per * kStep;
    if (p0 >= b1) continue;   // CTA-uniform
    const bool last = p1 >= b1;
    if (last) p1 = lim;
    stamp(P, 4);
    prefetch_piece(P, p0, p1);
    const long long slots = S.budget - (rank_base + chunk_prefix(P, S, chunk));  // left at the start of the chunk
    const long long cand = __ldcg(&P.ws->cand_cta[chunk]);
    stamp(P, 5);
    long long seen = 0;  // candidates in [b0, p1)
    if (slots <= 0 || slots >= cand) {  // nobody / everybody in this chunk gets a slot: no ranks needed
      const uint32_t grant = (cand != 0 && slots >= cand) ? UST_W_GRANTED : 0u;
      if (p1 > p0) {
        switch (variant) {
          case 3: uniform_range<true, true>(P, S, p0, p1, grant); break;
          case 2: uniform_range<true, false>(P, S, p0, p1, grant); break;
          case 1: uniform_range<false, true>(P, S, p0, p1, grant); break;
          default: uniform_range<false, false>(P, S, p0, p1, grant); break;
        }
      }
      if (last && lim < b1) {
        long long running = 0;
        general_step<false>(P, S, lim, b1, grant, running);
      }
      continue;
    }
    if (p1 > p0) {
      switch (variant) {
        case 3: seen = ordered_range<true, true, false>(P, S, p0, p1, slots, b0, false); break;
        case 2: seen = ordered_range<true, false, false>(P, S, p0, p1, slots, b0, false); break;
        case 1: seen = ordered_range<false, true, false>(P, S, p0, p1, slots, b0, false); break;
        default: seen = ordered_range<false, false, false>(P, S, p0, p1, slots, b0, false); break;
      }
    }
    stamp(P, 7);
    if (last && lim < b1) {  // ragged end of the whole array (< 128 nodes); p1 == lim
      long long running = 0;
      if (p1 > p0) {
        running = seen;
      } else {  // the piece is nothing but the ragged end: count what precedes it in the chunk
        long long before = 0;
        exact_prepass(P, S, p0, p0, b0, before);
        running = before;
      }
      if (threadIdx.x == 0) S.cand_prefix = rank_base + chunk_prefix(P, S, chunk);
      __syncthreads();
      general_step<true>(P, S, lim, b1, 0u, running);
      __syncthreads();
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.ws->fixups, (unsigned)m);
}

// After the grid barrier: redo what the streaming phase could not know. Aborts and pod lists: every CTA redoes
// its own chunks through the step path. Otherwise only a wrong slot speculation is left to repair.
__device__ void verify_phase(const UstParams& P, Shared& S, const ChunkCands& held) {
  const int n_chunks = P.grid_chunks;
  const long long rank_base = S.cand_prefix;  // candidates on lower ranks
  scan_chunks(P, S, rank_base, held);
  if (S.abort_key == ~0ull) {
    redo_wrong_chunks(P, S, rank_base);
    return;
  }
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const unsigned chunk_cand = __ldcg(&P.ws->cand_cta[chunk]);
    const long long before = chunk_prefix(P, S, chunk);
    __syncthreads();
    if (threadIdx.x == 0) S.cand_prefix = rank_base + before;
    __syncthreads();
    const long long b0 = chunk_bound(P.n, chunk, n_chunks), b1 = chunk_bound(P.n, chunk + 1, n_chunks);
    general_chunk(P, S, b0, b1, chunk_cand);
    if (threadIdx.x == 0) atomicAdd(&P.ws->fixups, 1u);
  }
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, UST_MIN_CTAS) ust_fused_kernel(const __grid_constant__ UstParams P) {
  __shared__ Shared S;
  stamp(P, 0);
  stage_tables_begin(P, S);
  __syncthreads();
  stream_phase(P, S);
  stamp(P, 1);
  // grid-wide barrier (every CTA is co-resident: cooperative launch). After it the cluster-wide
  // counters are final and the speculation can be checked.
  __syncthreads();
  ChunkCands held;
  if (P.fused_exchange) {
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&P.ws->arrive, 1u);
    }
    fused_exchange(P, S);
    stamp(P, 2);
    held = load_chunk_cands(P);
  } else {
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&P.ws->arrive, 1u);
      while (ld_acquire_u32(&P.ws->arrive) < gridDim.x) __nanosleep(20);
      __threadfence();
    }
    __syncthreads();
    stamp(P, 2);
    held = load_chunk_cands(P);  // same round trip as the vector
    load_local_vector(P, S);
    __syncthreads();
  }
  if (threadIdx.x == 0) derive_scalars(P, S);
  __syncthreads();
  if (verification_needed(P, S)) {
    verify_phase(P, S, held);
  }
  finish(P, S, true);
  stamp(P, 3);
}

// split mode (multi-GPU with a host-launched collective between the kernels): streaming phase ...
__global__ void __launch_bounds__(kThreads, UST_MIN_CTAS) ust_phase1_kernel(const __grid_constant__ UstParams P) {
  __shared__ Shared S;
  stage_tables_begin(P, S);
  __syncthreads();
  stream_phase(P, S);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    S.last = atomicAdd(&P.ws->depart, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (S.last) {  // last CTA of this launch
    __threadfence();
    if (P.publish) {  // ... and last streaming launch of the call: publish this shard's lanes, restore the invariant
      load_local_vector(P, S);
      __syncthreads();
      if (threadIdx.x < UST_V_LEN) P.xchg[threadIdx.x] = S.V[threadIdx.x];
      __syncthreads();
      if (threadIdx.x < 18) P.ws->acc[threadIdx.x] = 0;
      if (threadIdx.x == 0) P.ws->errinv = 0;
    }
    if (threadIdx.x == 0) { P.ws->depart = 0; P.ws->ticket = 0; }
  }
}

// ... and verification: redo, exactly, the chunks whose speculation did not hold
__global__ void __launch_bounds__(kThreads, UST_MIN_CTAS) ust_phase2_kernel(const __grid_constant__ UstParams P) {
  __shared__ Shared S;
  stage_tables_begin(P, S);
  if (threadIdx.x < UST_V_LEN) S.V[threadIdx.x] = P.xchg[threadIdx.x];
  const ChunkCands held = load_chunk_cands(P);
  __syncthreads();
  if (threadIdx.x == 0) derive_scalars(P, S);
  __syncthreads();
  stage_tables_wait(S);
  __syncthreads();
  if (verification_needed(P, S)) verify_phase(P, S, held);
  finish(P, S, false);
}

// Pod-list summaries (rows 12-14 of the scope table: pod_manager.go:256-391, :122-229, drain_manager.go:58-139).
// Only nodes whose actuator would look at its pods have their list read: wait-for-jobs, pod-deletion and
// drain-required nodes - everything else costs the hot byte. A CTA takes blocks of kPodBlock consecutive nodes:
// it compacts the nodes that need their list (in node order) into shared memory, then every thread walks the list
// of one such node with aligned 16-byte loads (8 pods each, all loads of a pass in flight together), mapping each
// pod through the per-policy pod table (shared memory) and OR-ing. All lanes of a warp do useful work on every
// instruction, which is what makes this an HBM-bound kernel instead of an issue-bound one (round-1 measurement:
// the warp-per-node formulation executed 15x the instructions). Neighbouring threads own neighbouring lists, so
// their loads share sectors. Output: one byte per node for the streaming pass (layout: pods_apply()), written
// coalesced per block.
constexpr int kPodBlock = 4096;            // nodes per block = 16 per thread
constexpr int kPodChunks = 6;              // 16-byte loads in flight per thread and pass (48 pods: a typical list in one pass)

__global__ void __launch_bounds__(kThreads, 6) ust_pod_summary_kernel(long long n, int active, const uint8_t* __restrict__ hot,
                                                                      const int32_t* __restrict__ pod_off,
                                                                      const uint16_t* __restrict__ pod_flags, long long n_pods,
                                                                      const uint8_t* __restrict__ podlut, uint8_t* __restrict__ podsum) {
  __shared__ __align__(16) uint8_t lut[UST_PODLUT_ENTRIES];
  __shared__ __align__(16) uint8_t res[kPodBlock];   // summary byte per node of the block (bits 6-7: state - 3 while in work)
  __shared__ unsigned short list[kPodBlock];         // block-local indices of the nodes whose list is read
  __shared__ int cnt;
  const int t = threadIdx.x, lane = t & 31;
  for (int i = t; i < (int)(UST_PODLUT_ENTRIES / 4); i += kThreads)
    reinterpret_cast<uint32_t*>(lut)[i] = __ldg(reinterpret_cast<const uint32_t*>(podlut) + i);
  const unsigned char* bytes = reinterpret_cast<const unsigned char*>(pod_flags);
  const long long total_bytes = 2 * n_pods;
  // hot bytes of this thread's 16 nodes of a block ("excluded" past the end of the array)
  auto load_hot = [&](long long base) -> uint4 {
    const long long i0 = base + 16 * t;
    uint32_t w[4] = {0x0E0E0E0Eu, 0x0E0E0E0Eu, 0x0E0E0E0Eu, 0x0E0E0E0Eu};
    if (i0 + 16 <= n) return __ldg(reinterpret_cast<const uint4*>(hot + i0));  // base, 16t: multiples of 16
    for (int k = 0; k < 16; k++)
      if (i0 + k < n) w[k >> 2] = (w[k >> 2] & ~(0xFFu << (8 * (k & 3)))) | ((uint32_t)hot[i0 + k] << (8 * (k & 3)));
    return make_uint4(w[0], w[1], w[2], w[3]);
  };
  for (long long base = (long long)blockIdx.x * kPodBlock; base < n; base += (long long)gridDim.x * kPodBlock) {
    if (t == 0) cnt = 0;
    __syncthreads();
    // ---- scan the block's hot bytes, compact the nodes that need their pods
    const long long i0 = base + 16 * t;
    const uint4 hv = load_hot(base);
    const uint32_t w[4] = {hv.x, hv.y, hv.z, hv.w};
    unsigned needbits = 0;
    uint32_t init[4];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const unsigned s = (w[k >> 2] >> (8 * (k & 3))) & 15u;
      const bool need = active && s >= UST_STATE_WAIT_FOR_JOBS_REQUIRED && s <= UST_STATE_DRAIN_REQUIRED;
      // wait-for-jobs: the list overrides the pre-evaluated bit (0x10) even when it is empty
      const uint32_t b = need ? (((s - UST_STATE_WAIT_FOR_JOBS_REQUIRED) << 6) | (s == UST_STATE_WAIT_FOR_JOBS_REQUIRED ? 0x10u : 0u)) : 0u;
      if ((k & 3) == 0) init[k >> 2] = 0;
      init[k >> 2] |= b << (8 * (k & 3));
      needbits |= (need ? 1u : 0u) << k;
    }
    *reinterpret_cast<uint4*>(res + 16 * t) = make_uint4(init[0], init[1], init[2], init[3]);
    const int mine_cnt = __popc(needbits);
    int incl = mine_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    int wbase = 0;
    if (lane == 31) wbase = atomicAdd(&cnt, incl);   // warps land in arrival order: the list is node-ordered within a warp
    wbase = __shfl_sync(kFull, wbase, 31);
    int slot = wbase + incl - mine_cnt;
    while (needbits) {
      const int k = __ffs(needbits) - 1;
      needbits &= needbits - 1;
      list[slot++] = (unsigned short)(16 * t + k);
    }
    __syncthreads();
    // ---- one thread per listed node
    const int total = cnt;
    for (int q = t; q < total; q += kThreads) {
      const int li = list[q];
      const long long i = base + li;
      const int p0 = __ldg(pod_off + i), p1 = __ldg(pod_off + i + 1);
      const int len = p1 - p0;
      unsigned r = 0;
      // 16-byte chunks covering the list; the last one may reach past the list but, when `safe`, not past the array
      const long long c0 = (2LL * p0) & ~15LL;
      const int nchunks = len > 0 ? (int)((2LL * p1 - c0 + 15) >> 4) : 0;
      const bool safe = c0 + 16LL * nchunks <= total_bytes;
      if (safe) {
        const uint4* src = reinterpret_cast<const uint4*>(bytes + c0);
        int rel = (int)((c0 >> 1) - p0);  // pod index of the chunk's element 0, relative to the list (<= 0 for chunk 0)
        for (int cb = 0; cb < nchunks; cb += kPodChunks) {
          uint4 x[kPodChunks];
#pragma unroll
          for (int u = 0; u < kPodChunks; u++) x[u] = cb + u < nchunks ? __ldcs(src + cb + u) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int u = 0; u < kPodChunks; u++) {
            if (cb + u < nchunks) {
              const uint32_t wv[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
              for (int e = 0; e < 8; e++) {
                const uint32_t f = (e & 1) ? (wv[e >> 1] >> 16) & (UST_PODLUT_ENTRIES - 1) : wv[e >> 1] & (UST_PODLUT_ENTRIES - 1);
                if ((unsigned)(rel + e) < (unsigned)len) r |= lut[f];
              }
            }
            rel += 8;
          }
        }
      } else {  // the list ends within the last 16 bytes of the whole array: plain 2-byte loads
        for (int p = p0; p < p1; p++) r |= lut[__ldg(pod_flags + p) & (UST_PODLUT_ENTRIES - 1)];
      }
      const unsigned tag = res[li];
      const unsigned s = UST_STATE_WAIT_FOR_JOBS_REQUIRED + (tag >> 6);
      unsigned ps;
      if (s == UST_STATE_WAIT_FOR_JOBS_REQUIRED) ps = 0x10u | (r & UST_PODSUM_WAIT_RUNNING);
      else if (s == UST_STATE_POD_DELETION_REQUIRED) ps = r & (UST_PODSUM_TO_DELETE | UST_PODSUM_CANNOT_DELETE);
      else ps = r & UST_PODSUM_DRAIN_ERROR;
      res[li] = (uint8_t)ps;
    }
    __syncthreads();
    // ---- coalesced write of the block's bytes
    if (i0 + 16 <= n) {
      *reinterpret_cast<uint4*>(podsum + i0) = *reinterpret_cast<const uint4*>(res + 16 * t);
    } else {
      for (int k = 0; k < 16; k++)
        if (i0 + k < n) podsum[i0 + k] = res[16 * t + k];
    }
  }
}

// BuildState at wire level: the owner join itself. One entry per driver pod with the 128-bit UID of
// OwnerReferences[0] ((0, 0) = no owner reference: an orphaned pod, common_manager.go:225-227); the driver
// DaemonSets' UIDs arrive sorted with their original indices. A pod whose owner is none of them is dropped
// (GetPodsOwnedbyDs skips it, GetOrphanedPods does not take it: common_manager.go:190-222) - it gets ds_idx -2 and
// counts as "not in snapshot". 17 B read + 4 B written per pod. The DaemonSet map (common_manager.go:181-185, keyed
// by UID) is an open-addressing hash table built by the host at load factor <= 1/4 (ust_uid_hash, linear probing,
// (0, 0) = empty slot) and copied to shared memory: one 16-byte lookup per pod in the common case. Counting is
// byte-sliced as in the streaming pass; per-DaemonSet counts are packed byte counters (<= 8 DaemonSets) or
// warp-aggregated atomics.
constexpr int kUidTabSmem = 2048;  // hash slots held in shared memory (DaemonSets <= 512); larger tables stay in global memory

// UID = false is the index form (ust_build_state: the host has already resolved the owner, ds_idx_in holds it; an index
// outside [0, n_ds) counts for no DaemonSet): same counting machinery, no join, nothing dropped.
template <bool UID>
__global__ void __launch_bounds__(kThreads) ust_build_state_uid_kernel(long long n, const uint8_t* __restrict__ hot,
                                                                       const ulonglong2* __restrict__ owner,
                                                                       const int32_t* __restrict__ ds_idx_in, int n_ds,
                                                                       const ulonglong2* __restrict__ ds_tab,
                                                                       const int32_t* __restrict__ ds_tab_idx, int tab_slots,
                                                                       int32_t* __restrict__ ds_idx_out,
                                                                       unsigned long long* ds_count, UstWorkspace* ws) {
  __shared__ ulonglong2 tab[kUidTabSmem];
  __shared__ int ord[kUidTabSmem];
  __shared__ unsigned int cnt_ds[kUidTabSmem / 4];
  __shared__ unsigned long long inc[256];  // hot byte -> sixteen 4-bit one-hot increments (fields as in the streaming pass)
  __shared__ unsigned int cnt[16];
  const int t = threadIdx.x;
  const bool in_smem = UID ? tab_slots <= kUidTabSmem : n_ds <= kUidTabSmem / 4;  // UID: then n_ds <= kUidTabSmem / 4 too
  if (in_smem) {
    if (UID)
      for (int i = t; i < tab_slots; i += kThreads) { tab[i] = ds_tab[i]; ord[i] = ds_tab_idx[i]; }
    for (int i = t; i < n_ds; i += kThreads) cnt_ds[i] = 0;
  }
  {
    const unsigned b = t, code = b & 15u;
    unsigned long long v = 0;
    if (code < 14) {
      v = 1ull << (4 * code);
      if (b & (UST_HOT_UNSCHEDULABLE | UST_HOT_NOT_READY)) v |= 1ull << 56;
      if (code == UST_STATE_UPGRADE_REQUIRED && !(b & UST_HOT_SKIP)) v |= 1ull << 60;
    }
    inc[b] = v;
  }
  if (t < 16) cnt[t] = 0;
  __syncthreads();
  const ulonglong2* table = in_smem ? tab : ds_tab;
  const int* order = in_smem ? ord : ds_tab_idx;
  const unsigned slot_mask = (unsigned)tab_slots - 1u;  // tab_slots is a power of two
  uint32_t B[4] = {0, 0, 0, 0}, lo = 0, hi = 0;
  int pending = 0;
  long long excluded = 0;
  unsigned long long dsl = 0;  // n_ds <= 8: this thread's owned-pod count per DaemonSet, one byte each
  auto spill = [&]() {
    widen(lo, hi, B);
#pragma unroll
    for (int f = 0; f < 16; f++) {
      const unsigned v = p1_field(B, f);
      if (v) atomicAdd(&cnt[f], v);
    }
    if (dsl) {
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const unsigned v = (unsigned)(dsl >> (8 * q)) & 0xFFu;
        if (v) atomicAdd(&cnt_ds[q], v);
      }
      dsl = 0;
    }
    B[0] = B[1] = B[2] = B[3] = 0;
    pending = 0;
  };
  constexpr int kU = 4;  // pods per thread and iteration: four 16-byte loads in flight
  const long long stride = (long long)gridDim.x * kThreads * kU;
  for (long long i0 = (long long)blockIdx.x * kThreads * kU; i0 < n; i0 += stride) {  // warp-uniform trip count
    ulonglong2 u[kU];
    int dk[kU];
    unsigned hb[kU];
#pragma unroll
    for (int k = 0; k < kU; k++) {
      const long long i = i0 + (long long)k * kThreads + t;
      u[k] = make_ulonglong2(0ull, 0ull);
      dk[k] = -1;
      hb[k] = UST_STATE_EXCLUDED;
      if (i < n) {
        if (UID) u[k] = __ldcs(owner + i); else dk[k] = __ldcs(ds_idx_in + i);
        hb[k] = __ldg(hot + i);
      }
    }
#pragma unroll
    for (int k = 0; k < kU; k++) {
      const long long i = i0 + (long long)k * kThreads + t;
      const bool valid = i < n;
      int d = -2;
      if (valid) {
        if (!UID) {
          d = (dk[k] >= 0 && dk[k] < n_ds) ? dk[k] : -1;
        } else if ((u[k].x | u[k].y) == 0ull) {
          d = -1;  // IsOrphanedPod
        } else {
          unsigned slot = ust_uid_hash(u[k].x, u[k].y) & slot_mask;
          for (;;) {  // linear probing; the table is at most a quarter full
            const ulonglong2 e = table[slot];
            if (e.x == u[k].x && e.y == u[k].y) { d = order[slot]; break; }
            if ((e.x | e.y) == 0ull) break;  // empty slot: not a driver DaemonSet's pod
            slot = (slot + 1u) & slot_mask;
          }
        }
        if (UID) __stcs(ds_idx_out + i, d);
        if (d != -2 && (hb[k] & 15u) < 14u) {  // in the snapshot (the host marks a pending-unscheduled pod with code 14)
          const unsigned long long v = inc[hb[k]];
          lo += (uint32_t)v;
          hi += (uint32_t)(v >> 32);
        } else {
          excluded++;
        }
        if ((++pending & 7) == 0) widen(lo, hi, B);
      }
      // per-DaemonSet owned-pod counts (before the pending-skip, upgrade_state.go:128). A handful of DaemonSets
      // (the usual case): eight byte counters packed in a register, flushed with the other counters; otherwise
      // one atomic per distinct DaemonSet per warp
      if (n_ds <= 8) {
        if (d >= 0) dsl += 1ull << (8 * d);
      } else {
        const unsigned act = __ballot_sync(kFull, d >= 0);
        if (d >= 0) {
          const unsigned peers = __match_any_sync(act, d);
          if ((t & 31) == __ffs(peers) - 1) {
            if (in_smem) atomicAdd(&cnt_ds[d], (unsigned)__popc(peers));
            else atomicAdd(&ds_count[d], (unsigned long long)__popc(peers));
          }
        }
      }
    }
    if (pending >= 240) spill();
  }
  spill();
  __syncthreads();
  // fields 0..13 per state code, 14 unavailable, 15 candidates -> ws->acc[0..13], [16], [17]
  for (int o = 16; o > 0; o >>= 1) excluded += __shfl_xor_sync(kFull, excluded, o);
  if ((t & 31) == 0 && excluded) atomicAdd(&ws->acc[UST_STATE_EXCLUDED], (unsigned long long)excluded);
  if (t < 14) { if (cnt[t]) atomicAdd(&ws->acc[t], (unsigned long long)cnt[t]); }
  else if (t == 14) { if (cnt[14]) atomicAdd(&ws->acc[16], (unsigned long long)cnt[14]); }
  else if (t == 15) { if (cnt[15]) atomicAdd(&ws->acc[17], (unsigned long long)cnt[15]); }
  if (in_smem)
    for (int i = t; i < n_ds; i += kThreads)
      if (cnt_ds[i]) atomicAdd(&ds_count[i], (unsigned long long)cnt_ds[i]);
}

__global__ void ust_build_state_finish_kernel(int n_ds, const int32_t* ds_desired, unsigned long long* ds_count,
                                              UstWorkspace* ws, ust_counters* out) {
  ust_counters c;
  for (int i = 0; i < 16; i++) c.hist[i] = (long long)ws->acc[i];
  c.unavailable = (long long)ws->acc[16];
  c.candidates = (long long)ws->acc[17];
  c.total_managed = c.hist[0] + c.hist[1] + c.hist[2] + c.hist[3] + c.hist[4] + c.hist[5] + c.hist[8] + c.hist[9] +
                    c.hist[10] + c.hist[11] + c.hist[12];
  c.in_progress = c.total_managed - c.hist[0] - c.hist[11] - c.hist[1];
  c.max_unavailable = 0;
  c.upgrades_available = 0;
  c.error_code = UST_OK;
  c.error_index = -1;
  c.error_pass = -1;
  for (int d = 0; d < n_ds; d++)
    if ((unsigned long long)(long long)ds_desired[d] != ds_count[d]) {  // upgrade_state.go:128-131
      c.error_code = UST_ERR_DS_UNSCHEDULED;
      c.error_index = d;
      break;
    }
  for (int i = 0; i < 7; i++) c.reserved[i] = 0;
  *out = c;
  for (int i = 0; i < 18; i++) ws->acc[i] = 0;
  for (int d = 0; d < n_ds; d++) ds_count[d] = 0;
}

// Delta update of the resident snapshot: scatter the re-encoded nodes into the SoA arrays.
__global__ void __launch_bounds__(kThreads) ust_patch_kernel(long long m, const long long* __restrict__ idx,
                                                             const uint8_t* __restrict__ state, const uint32_t* __restrict__ flags,
                                                             const int32_t* __restrict__ pod_rev, const int32_t* __restrict__ ds_idx,
                                                             uint8_t* hot_out, uint32_t* flags_out, int32_t* rev_out, int32_t* ds_out) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long k = (long long)blockIdx.x * kThreads + threadIdx.x; k < m; k += stride) {
    const long long i = __ldg(idx + k);
    hot_out[i] = __ldg(state + k);
    flags_out[i] = __ldg(flags + k);
    rev_out[i] = __ldg(pod_rev + k);
    ds_out[i] = __ldg(ds_idx + k);
  }
}

// Rollout simulation: the state feedback between two reconciles under "ideal actuators" - every call
// the reference makes through its providers takes effect, every asynchronous actuator succeeds, and whatever a node
// is waiting for (jobs, pod readiness, validation) has happened by the next reconcile. One streaming pass, in place:
// 13 B read + up to 9 B written per node.
//   state   <- actuator_outcome when the pass scheduled an asynchronous actuator (pod_manager.go:393-403,
//              drain_manager.go:111-139), else next_state (NodeUpgradeStateProvider.ChangeNodeUpgradeState)
//   annotations per action bit (ChangeNodeUpgradeAnnotation, upgrade_suit_test.go:121-130)
//   CORDON / UNCORDON -> Spec.Unschedulable (cordon_manager.go:40-47)
//   RESTART_DRIVER_POD -> the DaemonSet controller recreates the pod at the current revision and it becomes ready;
//              an orphaned pod is not recreated: the node leaves the snapshot (no driver pod to list)
//   still waiting after the pass: wait-for-jobs with running pods -> the jobs finish; pod-restart with a synced pod
//              that is not ready -> it becomes ready; validation-required -> the validation pod becomes ready
__global__ void __launch_bounds__(kThreads) ust_feedback_kernel(long long n, uint8_t* hot, uint32_t* flags, int32_t* pod_rev,
                                                                const int32_t* __restrict__ ds_idx, int n_ds,
                                                                const int32_t* __restrict__ ds_rev,
                                                                const uint8_t* __restrict__ next, const uint16_t* __restrict__ actions,
                                                                const uint8_t* __restrict__ outcome, const ust_counters* step) {
  if (step->error_code != UST_OK) return;  // the reconcile returned an error: nothing it decided is fed back
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    unsigned b = hot[i];
    const unsigned s = b & 15u;
    if (s >= UST_STATE_OTHER) continue;  // other label values / not in the snapshot: never processed
    uint32_t f = flags[i];
    const unsigned a = actions[i];
    const unsigned oc = outcome[i];
    unsigned ns = next[i];
    if ((a & (UST_A_SCHEDULE_WAIT_CHECK | UST_A_SCHEDULE_POD_EVICTION | UST_A_SCHEDULE_DRAIN)) && oc != UST_OUTCOME_NONE) ns = oc;
    if (a & UST_A_CLEAR_UPGRADE_REQUESTED) f &= ~UST_F_UPGRADE_REQUESTED;
    if (a & UST_A_SET_INITIAL_STATE_ANNO) f |= UST_F_INITIAL_STATE_ANNO;
    if (a & UST_A_CLEAR_INITIAL_STATE_ANNO) f &= ~UST_F_INITIAL_STATE_ANNO;
    if (a & UST_A_CORDON) b |= UST_HOT_UNSCHEDULABLE;
    if (a & UST_A_UNCORDON) b &= ~UST_HOT_UNSCHEDULABLE;
    if (a & UST_A_UNBLOCK_SAFE_LOAD) f &= ~UST_F_SAFE_LOAD;
    if (a & UST_A_SET_WAIT_START) f |= UST_F_WAIT_START_ANNO;
    if (a & UST_A_CLEAR_WAIT_START) f &= ~(UST_F_WAIT_START_ANNO | UST_F_WAIT_TIMED_OUT | UST_F_WAIT_START_INVALID);
    int rev = pod_rev[i];
    if (a & UST_A_RESTART_DRIVER_POD) {
      const int d = ds_idx[i];
      if ((f & UST_F_POD_ORPHANED) || d < 0 || d >= n_ds) {
        ns = UST_STATE_EXCLUDED;
      } else {
        rev = ds_rev[d];
        f = (f | UST_F_POD_READY) & ~(UST_F_POD_FAILING | UST_F_POD_TERMINATING);
      }
    }
    if (ns == UST_STATE_WAIT_FOR_JOBS_REQUIRED) f &= ~UST_F_WAIT_PODS_RUNNING;
    if (ns == UST_STATE_POD_RESTART_REQUIRED) {
      f &= ~UST_F_POD_TERMINATING;
      if (!(f & UST_F_POD_FAILING)) f |= UST_F_POD_READY;
    }
    if (ns == UST_STATE_VALIDATION_REQUIRED) f |= UST_F_VALIDATION_DONE;
    hot[i] = (uint8_t)((b & 0xF0u) | (ns & 15u));
    flags[i] = f;
    pod_rev[i] = rev;
  }
}

// Packed host format (ust_apply_state_packed): the interned pod revision travels as uint16 and the DaemonSet index
// as int8 over PCIe; this widens a range of them into the int32 arrays the streaming pass reads. 3 B read + 8 B
// written per node, once per upload segment.
__global__ void __launch_bounds__(kThreads) ust_widen_kernel(long long n, const uint16_t* __restrict__ rev16,
                                                             const int8_t* __restrict__ ds8, int32_t* __restrict__ rev_out,
                                                             int32_t* __restrict__ ds_out) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    rev_out[i] = (int32_t)__ldcs(rev16 + i);
    ds_out[i] = (int32_t)__ldcs(ds8 + i);
  }
}

}  // namespace

int ust_launch_fused(const UstParams& p, int grid, void* stream) {
  void* args[] = {(void*)&p};
  return (int)cudaLaunchCooperativeKernel((const void*)ust_fused_kernel, dim3(grid), dim3(kThreads), args, 0, (cudaStream_t)stream);
}
int ust_launch_phase1(const UstParams& p, int grid, void* stream) {
  ust_phase1_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}
int ust_launch_phase2(const UstParams& p, int grid, void* stream) {
  ust_phase2_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}
int ust_launch_pod_summary(long long n, int active, const uint8_t* hot, const int32_t* pod_off, const uint16_t* pod_flags,
                           long long n_pods, const uint8_t* podlut, uint8_t* podsum, int grid, void* stream) {
  if (n <= 0) return 0;
  const long long blocks = (n + kPodBlock - 1) / kPodBlock;
  if (grid > blocks) grid = (int)blocks;
  ust_pod_summary_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, active, hot, pod_off, pod_flags, n_pods, podlut, podsum);
  return (int)cudaGetLastError();
}
int ust_launch_build_state(long long n, const uint8_t* hot, const int32_t* ds_idx, int n_ds, const int32_t* ds_desired,
                           unsigned long long* ds_count, UstWorkspace* ws, ust_counters* out, int grid, void* stream) {
  ust_build_state_uid_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, hot, nullptr, ds_idx, n_ds, nullptr, nullptr, 8,
                                                                                 nullptr, ds_count, ws);
  ust_build_state_finish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(n_ds, ds_desired, ds_count, ws, out);
  return (int)cudaGetLastError();
}
int ust_launch_widen(long long n, const uint16_t* rev16, const int8_t* ds8, int32_t* rev_out, int32_t* ds_out, int grid,
                     void* stream) {
  if (n <= 0) return 0;
  const long long want = (n + kThreads - 1) / kThreads;
  ust_widen_kernel<<<(unsigned)(want < grid ? want : grid), kThreads, 0, (cudaStream_t)stream>>>(n, rev16, ds8, rev_out, ds_out);
  return (int)cudaGetLastError();
}
int ust_launch_patch(long long m, const long long* idx, const uint8_t* state, const uint32_t* flags, const int32_t* pod_rev,
                     const int32_t* ds_idx, uint8_t* hot_out, uint32_t* flags_out, int32_t* rev_out, int32_t* ds_out, void* stream) {
  if (m <= 0) return 0;
  const long long grid = (m + kThreads - 1) / kThreads;
  ust_patch_kernel<<<(unsigned)(grid > 65535 * 16 ? 65535 * 16 : grid), kThreads, 0, (cudaStream_t)stream>>>(
      m, idx, state, flags, pod_rev, ds_idx, hot_out, flags_out, rev_out, ds_out);
  return (int)cudaGetLastError();
}
int ust_launch_feedback(long long n, uint8_t* hot, uint32_t* flags, int32_t* pod_rev, const int32_t* ds_idx, int n_ds,
                        const int32_t* ds_rev, const uint8_t* next, const uint16_t* actions, const uint8_t* outcome,
                        const ust_counters* step, int grid, void* stream) {
  if (n <= 0) return 0;
  ust_feedback_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, hot, flags, pod_rev, ds_idx, n_ds, ds_rev, next, actions, outcome, step);
  return (int)cudaGetLastError();
}
int ust_launch_build_state_uids(long long n, const uint8_t* hot, const void* owner_uid, int n_ds, const void* ds_tab,
                                const int32_t* ds_tab_idx, int tab_slots, const int32_t* ds_desired, int32_t* ds_idx_out,
                                unsigned long long* ds_count, UstWorkspace* ws, ust_counters* out, int grid, void* stream) {
  ust_build_state_uid_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
      n, hot, reinterpret_cast<const ulonglong2*>(owner_uid), nullptr, n_ds, reinterpret_cast<const ulonglong2*>(ds_tab), ds_tab_idx,
      tab_slots, ds_idx_out, ds_count, ws);
  ust_build_state_finish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(n_ds, ds_desired, ds_count, ws, out);
  return (int)cudaGetLastError();
}
int ust_max_coresident_ctas(int device, int* ctas_per_sm, int* num_sms) {
  int per_sm = 0, sms = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ust_fused_kernel, kThreads, 0);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return (int)e;
  *ctas_per_sm = per_sm;
  *num_sms = sms;
  return 0;
}
