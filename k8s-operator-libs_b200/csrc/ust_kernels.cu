// ust_kernels.cu — the ApplyState kernels for sm_100a (B200).
//
// What replaces what: one launch of ust_fused_kernel computes, for every node of the snapshot, what the
// reference's ClusterUpgradeStateManagerImpl.ApplyState (pkg/upgrade/upgrade_state.go:171-281) computes
// with its twelve sequential Process* loops: next state label and actuator-call bitmask per node, plus
// the cluster counters of common_manager.go:715-788.
//
// Shape of the kernel (HBM-bound byte/integer streaming, no tensor-core work):
//   * persistent, cooperative grid of (#SM x resident CTAs); CTA c owns one contiguous chunk of nodes, so
//     slice order of the upgrade-required bucket (upgrade_inplace.go:71) is chunk order;
//   * phase 1 streams ONLY the 1-byte hot array (state code + 4 constraint predicates): 14-bin state
//     histogram, unavailable count, upgrade candidates per chunk, earliest abort point — byte-sliced
//     SIMD-in-register counters fed from a 256-entry shared-memory table, one REDUX per counter per warp,
//     16 global atomics per CTA;
//   * one grid-wide barrier (single global atomic counter); every CTA then derives the slot budget
//     (GetUpgradesAvailable, common_manager.go:748-776) and its exclusive candidate prefix;
//   * phase 2 streams state(1 B, L2-resident by now) + flags(4) + pod_rev(4) + ds_idx(4) with 128-bit
//     coalesced loads, evaluates each node by ONE shared-memory table lookup (policy staged in shared
//     memory as a per-state transition table, see ust_lut.h) and writes next_state(1) + actions(2) with
//     full-width coalesced stores: 16 algorithmic bytes per node, each touched once;
//   * only the single chunk that straddles the slot budget runs the exact ordered path (warp-shuffle +
//     shared-memory exclusive scan of candidate bits); all other chunks are on one side of the cut.
#include <cuda_runtime.h>

#include "ust_dev.h"

namespace {

constexpr int kThreads = UST_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kStep = kThreads * 4;   // nodes per CTA step in phase 2 (4 per thread)
#ifndef UST_UNROLL
#define UST_UNROLL 4
#endif
#ifndef UST_DOUBLE_BUFFER
#define UST_DOUBLE_BUFFER 0
#endif
#ifndef UST_MIN_CTAS
#define UST_MIN_CTAS 2
#endif
constexpr int kUnroll = UST_UNROLL;            // steps in flight per thread in the fast path
constexpr int kTile = kStep * kUnroll;
constexpr int kPrefetchTiles = 2;     // L2 prefetch distance of the phase-2 stream, in tiles
constexpr int kP1Unroll = 4;          // 16-byte hot loads in flight per thread in phase 1
constexpr unsigned kFull = 0xFFFFFFFFu;
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
// nothing waits for it until phase 2 starts.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stamp(const UstParams& P, int k) {
  if (threadIdx.x == 0 && k < 4) {
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
  if (t == 32) { S.errinv = 0; S.abort_key = ~0ull; S.chunk_cand = 0; }
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
// phase 1: counts over the hot bytes of [b0, b1)
//
// Byte-sliced SIMD-in-register counting. A 256-entry shared-memory table maps a hot byte to sixteen
// 4-bit one-hot increments packed in 64 bits (fields 0-13: state code, 14: unavailable, 15: upgrade
// candidate); a thread sums the entries of 8 nodes (no field can exceed 8), widens the nibbles to byte
// lanes, and keeps going. Per node: one 8-byte LDS, two shifts/masks for the address, one add. No
// atomics until the end of the chunk: one REDUX per counter per warp, 16 global atomics per CTA.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void p1_error_byte(const UstParams& P, Shared& S, unsigned b, long long i) {
  const unsigned code = b & 15u;
  if (!(b & UST_HOT_REVISION_HASH_ERROR) || !P.active) return;
  if (!(code == UST_STATE_UNKNOWN || code == UST_STATE_DONE || code == UST_STATE_POD_RESTART_REQUIRED || code == UST_STATE_FAILED)) return;
  if (__ldg(P.flags + i) & UST_F_POD_ORPHANED) return;  // orphaned pods never reach the hash lookup (common_manager.go:301-303)
  const unsigned long long key = UST_KEY(pass_of_state(code), (unsigned long long)i + 1ull);
  atomicMax(&S.errinv, ~key);
}

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

__device__ __forceinline__ void p1_words(const Shared& S, uint32_t x, uint32_t y, uint32_t (&B)[4]) {
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint2 a = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(S.hotent) + (((x >> (8 * k)) & 0xFFu) << 4) + 8);
    const uint2 c = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(S.hotent) + (((y >> (8 * k)) & 0xFFu) << 4) + 8);
    lo += a.x + c.x;
    hi += a.y + c.y;
  }
  B[0] += lo & 0x0F0F0F0Fu;
  B[1] += (lo >> 4) & 0x0F0F0F0Fu;
  B[2] += hi & 0x0F0F0F0Fu;
  B[3] += (hi >> 4) & 0x0F0F0F0Fu;
}

__device__ void phase1_count(const UstParams& P, Shared& S, long long b0, long long b1) {
  const int t = threadIdx.x;
  uint32_t B[4] = {0, 0, 0, 0};
  int pending = 0;
  constexpr uint32_t kFill = 0x0E0E0E0Eu;  // "excluded": contributes to no counter
  for (long long base = b0; base < b1; base += 16LL * kThreads * kP1Unroll) {
    uint4 h[kP1Unroll];
#pragma unroll
    for (int u = 0; u < kP1Unroll; u++) {
      const long long i = base + (long long)u * 16 * kThreads + 16LL * t;
      if (i + 16 <= b1) {
        h[u] = __ldg(reinterpret_cast<const uint4*>(P.hot + i));
      } else {
        uint32_t w[4] = {kFill, kFill, kFill, kFill};
        for (long long j = i; j < b1; j++) {  // ragged end of the array (at most 15 bytes, one lane)
          const int q = (int)(j - i);
          w[q >> 2] = (w[q >> 2] & ~(0xFFu << (8 * (q & 3)))) | ((uint32_t)P.hot[j] << (8 * (q & 3)));
        }
        h[u] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
#pragma unroll
    for (int u = 0; u < kP1Unroll; u++) {
      if ((h[u].x | h[u].y | h[u].z | h[u].w) & 0x80808080u) {  // rare: a revision-hash error bit among these 16 nodes
        const uint32_t w[4] = {h[u].x, h[u].y, h[u].z, h[u].w};
        const long long i = base + (long long)u * 16 * kThreads + 16LL * t;
        for (int q = 0; q < 16; q++) {
          const unsigned b = (w[q >> 2] >> (8 * (q & 3))) & 0xFFu;
          if ((b & 15u) < UST_STATE_EXCLUDED) p1_error_byte(P, S, b, i + q);
        }
      }
      p1_words(S, h[u].x, h[u].y, B);
      p1_words(S, h[u].z, h[u].w, B);
    }
    pending += 16 * kP1Unroll;
    if (pending > 255 - 16 * kP1Unroll) {  // byte lanes hold at most 255: spill (thread-serial, rare)
#pragma unroll
      for (int f = 0; f < 16; f++) {
        const unsigned v = p1_field(B, f);
        if (v) atomicAdd(&S.cnt[f], v);
      }
      B[0] = B[1] = B[2] = B[3] = 0;
      pending = 0;
    }
  }
  // all threads converged: one REDUX per counter per warp, one shared atomic per warp
#pragma unroll
  for (int f = 0; f < 16; f++) {
    const unsigned v = __reduce_add_sync(kFull, p1_field(B, f));
    if ((t & 31) == 0 && v) atomicAdd(&S.cnt[f], v);
  }
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
// phase 2: per-node transition
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

// pod-list summary for one node (thread-serial; only nodes whose actuator would run are evaluated)
__device__ __forceinline__ uint32_t pod_summary(const UstParams& P, uint32_t hb, long long i, uint32_t& clear_mask) {
  const unsigned s = hb & 15u;
  clear_mask = 0;
  if (s < UST_STATE_WAIT_FOR_JOBS_REQUIRED || s > UST_STATE_DRAIN_REQUIRED) return 0;
  unsigned r = 0;
  const int p0 = __ldg(P.pod_off + i), p1 = __ldg(P.pod_off + i + 1);
  for (int p = p0; p < p1; p++) r |= __ldg(P.podlut + (__ldg(P.pod_flags + p) & (UST_PODLUT_ENTRIES - 1)));
  uint32_t extra = 0;
  if (s == UST_STATE_WAIT_FOR_JOBS_REQUIRED) {
    clear_mask = UST_F_WAIT_PODS_RUNNING;  // the pod list, when given, overrides the pre-evaluated bit
    if (r & UST_PODSUM_WAIT_RUNNING) extra |= UST_F_WAIT_PODS_RUNNING;
  } else if (s == UST_STATE_POD_DELETION_REQUIRED) {
    if (r & UST_PODSUM_TO_DELETE) extra |= UST_W_PD_HAS;
    if (r & UST_PODSUM_CANNOT_DELETE) extra |= UST_W_PD_MISMATCH;
  } else {
    if (r & UST_PODSUM_DRAIN_ERROR) extra |= UST_W_DRAIN_ERROR;
  }
  return extra;
}

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
struct Tile {
  uint32_t h[kUnroll];
  uint4 f[kUnroll], r[kUnroll], d[kUnroll];
};

// Chunk-relative addressing: the chunk's base pointers are CTA-uniform; a thread addresses 4-node groups
// with a 32-bit group index q (thread t of the CTA owns groups done/4 + j*kStepQ + t of a tile).
struct Cursor {
  const uint32_t* h;
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
  c.f = reinterpret_cast<const uint4*>(P.flags + base);
  c.r = reinterpret_cast<const uint4*>(P.pod_rev + base);
  c.d = reinterpret_cast<const uint4*>(P.ds_idx + base);
  c.nx = reinterpret_cast<uint32_t*>(P.next + base);
  c.ac = reinterpret_cast<uint2*>(P.actions + base);
  c.oc = P.outcome ? reinterpret_cast<uint32_t*>(P.outcome + base) : nullptr;
  c.q = threadIdx.x;
  return c;
}
__device__ __forceinline__ void cursor_advance(Cursor& c) { c.q += kTileQ; }

// `room` = nodes left in the chunk from this thread's first node of the tile; chunk ends are multiples of
// 128 nodes, so for a partial tile validity is uniform per warp and per step.
template <bool FULL>
__device__ __forceinline__ void tile_load(const Cursor& c, int room, Tile& T) {
#pragma unroll
  for (int j = 0; j < kUnroll; j++) {
    if (FULL || j * kStep + 4 <= room) {
      T.h[j] = __ldg(c.h + c.q + j * kStepQ);
      T.f[j] = __ldcs(c.f + c.q + j * kStepQ);
      T.r[j] = __ldcs(c.r + c.q + j * kStepQ);
      T.d[j] = __ldcs(c.d + c.q + j * kStepQ);
    }
  }
}

// pull the int32 arrays of the tile `ahead` tiles further on into L2 (one request per 128-byte line)
__device__ __forceinline__ void tile_prefetch_l2(const Cursor& c, int ahead, int room) {
  if ((threadIdx.x & 7) != 0) return;
#pragma unroll
  for (int j = 0; j < kUnroll; j++) {
    if (ahead * kTile + j * kStep + 4 <= room) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(c.f + c.q + ahead * kTileQ + j * kStepQ));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(c.r + c.q + ahead * kTileQ + j * kStepQ));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(c.d + c.q + ahead * kTileQ + j * kStepQ));
    }
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
template <bool DS_SMEM>
__device__ __forceinline__ uint32_t stream_node(const UstParams& P, const Shared& S, uint32_t tab_off, uint32_t wbits,
                                                uint32_t fl, int rev, uint32_t di, uint32_t grant, uint32_t& lo,
                                                uint32_t& hi) {
  const uint4 m = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(S.hotent) + tab_off);
  lo += m.z;
  hi += m.w;
  uint32_t w = (fl & UST_F_INPUT_MASK) | wbits | grant;
  if (pod_synced<DS_SMEM>(P, S, rev, di)) w |= UST_W_SYNCED;
  const uint32_t off = (__funnelshift_r(w, 0u, m.x) & 0x7FCu) | m.y;
  return *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(S.lut) + off);
}

// The single streaming pass: for one tile, count (byte-sliced) AND evaluate every node with the chunk's
// speculative slot grant, writing next_state / actions. Whether the speculation was right is only known
// after the grid barrier; chunks on the wrong side of the cut are redone exactly there.
template <bool FULL, bool DS_SMEM, bool OUTCOME>
__device__ __forceinline__ void spec_tile(const UstParams& P, Shared& S, const Cursor& c, int room, long long i0,
                                          const Tile& T, uint32_t grant, uint32_t (&B)[4]) {
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int j = 0; j < kUnroll; j++) {
    if (FULL || j * kStep + 4 <= room) {
      const uint32_t x = T.h[j];
      if (x & 0x80808080u) {  // rare
        const long long i = i0 + j * kStep;
        spec_error_byte(P, S, x & 0xFFu, T.f[j].x, i);
        spec_error_byte(P, S, (x >> 8) & 0xFFu, T.f[j].y, i + 1);
        spec_error_byte(P, S, (x >> 16) & 0xFFu, T.f[j].z, i + 2);
        spec_error_byte(P, S, x >> 24, T.f[j].w, i + 3);
      }
      constexpr uint32_t kW = UST_W_SKIP | UST_W_UNSCHEDULABLE;
      uint32_t e[4];
      e[0] = stream_node<DS_SMEM>(P, S, (x << 4) & 0xFF0u, (x >> 3) & kW, T.f[j].x, (int)T.r[j].x, T.d[j].x, grant, lo, hi);
      e[1] = stream_node<DS_SMEM>(P, S, (x >> 4) & 0xFF0u, (x >> 11) & kW, T.f[j].y, (int)T.r[j].y, T.d[j].y, grant, lo, hi);
      e[2] = stream_node<DS_SMEM>(P, S, (x >> 12) & 0xFF0u, (x >> 19) & kW, T.f[j].z, (int)T.r[j].z, T.d[j].z, grant, lo, hi);
      e[3] = stream_node<DS_SMEM>(P, S, (x >> 20) & 0xFF0u, (x >> 27) & kW, T.f[j].w, (int)T.r[j].w, T.d[j].w, grant, lo, hi);
      uint32_t next4, out4;
      uint2 act4;
      pack4(e, next4, act4, out4);
      __stcs(c.nx + c.q + j * kStepQ, next4);
      __stcs(c.ac + c.q + j * kStepQ, act4);
      if (OUTCOME) __stcs(c.oc + c.q + j * kStepQ, out4);
    }
    if (j & 1) widen(lo, hi, B);  // at most 8 per nibble so far
  }
  if (kUnroll & 1) widen(lo, hi, B);
}

// The chunk loop is software-pipelined over two register tiles: tile i+1's loads are issued before tile i
// is evaluated, so HBM always has a full tile per thread in flight while the SM computes.
template <bool DS_SMEM, bool OUTCOME>
__device__ __forceinline__ void spec_chunk(const UstParams& P, Shared& S, long long b0, long long lim, uint32_t grant, Acc& A,
                           bool wait_for_table) {
  const long long span = lim - b0;  // CTA-uniform, a multiple of 128
  const int t4 = 4 * threadIdx.x;
  auto room_at = [&](long long done) -> int {  // nodes from this thread's first node of the tile to the chunk end
    const long long r = span - done - t4;
    return r > (1LL << 30) ? (1 << 30) : (r < 0 ? 0 : (int)r);
  };
  Cursor c = cursor_at(P, b0);
  long long i0 = b0 + t4;
  Tile T;
  auto load = [&](long long done) {  // loads of the tile starting `done` nodes into the chunk; c.q points at it
    if (span - done >= kTile) tile_load<true>(c, 0, T); else tile_load<false>(c, room_at(done), T);
  };
  if (span > 0) load(0);  // the first tile's loads go out before anything waits on the table copy
  if (wait_for_table) {
    stage_tables_wait(S);
    __syncthreads();
  }
  for (long long done = 0; done < span; done += kTile) {
    if (done) load(done);
    const int room = room_at(done);
    if (span - done >= kTile) spec_tile<true, DS_SMEM, OUTCOME>(P, S, c, room, i0, T, grant, A.B);
    else spec_tile<false, DS_SMEM, OUTCOME>(P, S, c, room, i0, T, grant, A.B);
    cursor_advance(c);
    i0 += kTile;
    if (++A.tiles >= 14) spill_thread(S, A);  // byte lanes: at most 16 per tile, 255 max
  }
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
    if (P.eval_pods && k < nvalid) {
      uint32_t clear_mask;
      extra |= pod_summary(P, hb[k], i0 + k, clear_mask);
      f &= ~clear_mask;
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

// exact / abort / pod-list evaluation of a whole chunk (the verification step falls back to this)
__device__ void general_chunk(const UstParams& P, Shared& S, long long b0, long long b1, unsigned chunk_cand) {
  const long long lo = S.cand_prefix, hi = S.cand_prefix + chunk_cand;
  const bool slotted = P.active && !P.requestor;
  const bool exact = slotted && chunk_cand != 0 && lo < S.budget && hi > S.budget;
  const uint32_t grant = (slotted && chunk_cand != 0 && hi <= S.budget) ? UST_W_GRANTED : 0u;
  long long running = 0;
  if (!exact) {
    for (long long base = b0; base < b1; base += kStep) general_step<false>(P, S, base, b1, grant, running);
  } else {
    for (long long base = b0; base < b1; base += kStep) general_step<true>(P, S, base, b1, 0u, running);
  }
}

// speculative grant of a chunk: chunks before P.spec_cut_chunk assume every candidate gets a slot
__device__ __forceinline__ uint32_t spec_grant(const UstParams& P, int chunk) {
  return (P.active && !P.requestor && chunk < P.spec_cut_chunk) ? UST_W_GRANTED : 0u;
}

// was the speculation right for this chunk? (called after derive_scalars + prefix)
__device__ __forceinline__ bool spec_holds(const UstParams& P, const Shared& S, int chunk, unsigned chunk_cand) {
  if (S.abort_key != ~0ull) return false;
  if (!(P.active && !P.requestor) || chunk_cand == 0) return true;
  const long long lo = S.cand_prefix, hi = S.cand_prefix + chunk_cand;
  return spec_grant(P, chunk) ? (hi <= S.budget) : (lo >= S.budget);
}

// The chunk loop of the streaming phase: chunks (contiguous node ranges, chunk order == slice order) are
// claimed with an atomic ticket, so CTAs that HBM serves faster take more of them and all CTAs reach the
// grid barrier together. Per chunk: counts + speculative outputs; the chunk's candidate count is published
// for the ordered slot allocation. Returns the number of nodes this CTA streamed.
template <bool DS_SMEM, bool OUTCOME>
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
    const uint32_t grant = spec_grant(P, chunk);
    const unsigned cand0 = A.cand_spilled + p1_field(A.B, 15);
    spec_chunk<DS_SMEM, OUTCOME>(P, S, b0, lim, grant, A, first);
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
  const int n_chunks = P.grid_chunks;
  UstWorkspace* ws = P.ws;
  long long nodes_seen = 0;
  if (P.eval_pods) {
    // pod lists are evaluated by the exact path after the barrier: count only, one static chunk per CTA
    const int chunk = P.chunk_begin + blockIdx.x;
    if (chunk < P.chunk_end) {
      const long long b0 = chunk_bound(P.n, chunk, n_chunks), b1 = chunk_bound(P.n, chunk + 1, n_chunks);
      phase1_count(P, S, b0, b1);
      nodes_seen = b1 - b0;
      __syncthreads();
      if (t == 0) ws->cand_cta[chunk] = S.cnt[15];
    }
  } else if (P.n_ds <= UST_DS_SMEM_MAX) {
    nodes_seen = P.outcome ? stream_loop<true, true>(P, S) : stream_loop<true, false>(P, S);
  } else {
    nodes_seen = P.outcome ? stream_loop<false, true>(P, S) : stream_loop<false, false>(P, S);
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

__device__ long long block_sum_cand_before(const UstParams& P, Shared& S, int chunk) {
  // exclusive prefix of per-chunk candidate counts (chunk order == slice order)
  const int t = threadIdx.x;
  long long s = 0;
  for (int c = t; c < chunk; c += kThreads) s += __ldcg(&P.ws->cand_cta[c]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
  if ((t & 31) == 0) S.part[t >> 5] = s;
  __syncthreads();
  long long tot = 0;
  for (int w = 0; w < kWarps; w++) tot += S.part[w];
  __syncthreads();
  return tot;
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

__device__ void finish(const UstParams& P, Shared& S, bool reset_ws) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(&P.ws->depart, 1u);
    if (prev == gridDim.x - 1u) {  // last CTA out: publish counters, restore the workspace invariant
      write_counters(P, S);
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
  if (P.eval_pods || S.abort_key != ~0ull) return true;
  if (!(P.active && !P.requestor)) return false;
  const long long cands = S.V[UST_V_CANDIDATES];
  if (cands == 0) return false;
  return P.spec_cut_chunk ? (S.budget < cands) : (S.budget > 0);
}

// redo, exactly, every chunk of this CTA's share whose speculation did not hold (or that needs pod lists)
__device__ void verify_phase(const UstParams& P, Shared& S) {
  const int n_chunks = P.grid_chunks;
  const long long rank_base = S.cand_prefix;  // candidates on lower ranks
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const long long before = block_sum_cand_before(P, S, chunk);
    const unsigned chunk_cand = __ldcg(&P.ws->cand_cta[chunk]);
    if (threadIdx.x == 0) S.cand_prefix = rank_base + before;
    __syncthreads();
    if (P.eval_pods || !spec_holds(P, S, chunk, chunk_cand)) {
      const long long b0 = chunk_bound(P.n, chunk, n_chunks), b1 = chunk_bound(P.n, chunk + 1, n_chunks);
      general_chunk(P, S, b0, b1, chunk_cand);
      if (threadIdx.x == 0) atomicAdd(&P.ws->fixups, 1u);
    }
    __syncthreads();
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
  if (P.fused_exchange) {
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&P.ws->arrive, 1u);
    }
    fused_exchange(P, S);
    stamp(P, 2);
  } else {
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&P.ws->arrive, 1u);
      while (ld_acquire_u32(&P.ws->arrive) < gridDim.x) __nanosleep(20);
      __threadfence();
    }
    __syncthreads();
    stamp(P, 2);
    load_local_vector(P, S);
    __syncthreads();
  }
  if (threadIdx.x == 0) derive_scalars(P, S);
  __syncthreads();
  if (verification_needed(P, S)) {
    if (P.eval_pods) { stage_tables_wait(S); __syncthreads(); }
    verify_phase(P, S);
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
  if (P.eval_pods) stage_tables_wait(S);  // never leave a bulk copy in flight at CTA exit
}

// ... and verification: redo, exactly, the chunks whose speculation did not hold
__global__ void __launch_bounds__(kThreads, UST_MIN_CTAS) ust_phase2_kernel(const __grid_constant__ UstParams P) {
  __shared__ Shared S;
  stage_tables_begin(P, S);
  if (threadIdx.x < UST_V_LEN) S.V[threadIdx.x] = P.xchg[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) derive_scalars(P, S);
  __syncthreads();
  stage_tables_wait(S);
  __syncthreads();
  if (verification_needed(P, S)) verify_phase(P, S);
  finish(P, S, false);
}

// BuildState's device part (upgrade_state.go:126-133, :158-160): owned pods per DaemonSet and bucket sizes
__global__ void __launch_bounds__(kThreads) ust_build_state_kernel(long long n, const uint8_t* hot, const int32_t* ds_idx,
                                                                   int n_ds, unsigned long long* ds_count,
                                                                   UstWorkspace* ws) {
  __shared__ unsigned int cnt[18];
  if (threadIdx.x < 18) cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned b = hot[i];
    unsigned code = b & 15u;
    if (code == 15) code = 14;
    atomicAdd(&cnt[code], 1u);
    if (code < 14 && (b & (UST_HOT_UNSCHEDULABLE | UST_HOT_NOT_READY))) atomicAdd(&cnt[16], 1u);
    if (code == UST_STATE_UPGRADE_REQUIRED && !(b & UST_HOT_SKIP)) atomicAdd(&cnt[17], 1u);
    const int d = ds_idx[i];
    if (d >= 0 && d < n_ds) atomicAdd(&ds_count[d], 1ull);
  }
  __syncthreads();
  if (threadIdx.x < 18 && cnt[threadIdx.x]) atomicAdd(&ws->acc[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
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
int ust_launch_build_state(long long n, const uint8_t* hot, const int32_t* ds_idx, int n_ds, const int32_t* ds_desired,
                           unsigned long long* ds_count, UstWorkspace* ws, ust_counters* out, int grid, void* stream) {
  ust_build_state_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, hot, ds_idx, n_ds, ds_count, ws);
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
