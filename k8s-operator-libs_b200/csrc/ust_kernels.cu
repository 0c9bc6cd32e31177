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
constexpr int kStep = kThreads * 4;  // nodes per CTA step in phase 2 (4 per thread)
constexpr int kUnroll = 4;           // steps in flight per thread in the fast path

struct __align__(16) Shared {
  uint32_t lut[UST_LUT_ENTRIES];
  uint4 hotlut[256];
  uint2 meta[16];
  int dsrev[UST_DS_SMEM_MAX + 1];
  unsigned int cnt[16];
  unsigned long long errinv;
  long long V[UST_V_LEN];
  // derived, CTA-uniform
  unsigned long long abort_key;  // ~0 = none
  long long budget;              // max(upgradesAvailable, 0)
  long long avail;
  long long max_unav;
  long long node_offset;         // global index of this shard's node 0
  long long cand_prefix;         // candidates before this chunk (global order)
  unsigned int warp_tot[kWarps];
  unsigned int chunk_cand;
  int red_scratch[kWarps];
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
  return b & ~127LL;  // chunks start on 128-node boundaries: 128 B of hot bytes, 512 B of each int32 array
}

// ------------------------------------------------------------------------------------------------
// table staging
// ------------------------------------------------------------------------------------------------
__device__ void stage_tables(const UstParams& P, Shared& S) {
  const int t = threadIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(P.lut);
  uint4* dst = reinterpret_cast<uint4*>(S.lut);
  for (int i = t; i < (int)(UST_LUT_ENTRIES / 4); i += kThreads) dst[i] = __ldg(src + i);
  if (t < 16) S.meta[t] = __ldg(reinterpret_cast<const uint2*>(P.lut + UST_LUT_ENTRIES) + t);
  if (P.n_ds <= UST_DS_SMEM_MAX)
    for (int i = t; i <= P.n_ds; i += kThreads) S.dsrev[i] = i < P.n_ds ? __ldg(P.ds_rev + i) : 0;
  // hot-byte -> byte-sliced counter increments: field j<14 = (code==j), field 14 = unavailable
  // (GetCurrentUnavailableNodes, common_manager.go:146-165), field 15 = upgrade candidate
  // (upgrade-required and not skip, upgrade_inplace.go:82)
  {
    const unsigned b = t, code = b & 15u;
    unsigned w[4] = {0, 0, 0, 0};
    if (code < 14) {
      w[code >> 2] |= 1u << (8 * (code & 3));
      if (b & (UST_HOT_UNSCHEDULABLE | UST_HOT_NOT_READY)) w[3] |= 1u << 16;
      if (code == UST_STATE_UPGRADE_REQUIRED && !(b & UST_HOT_SKIP)) w[3] |= 1u << 24;
    }
    S.hotlut[b] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (t < 16) S.cnt[t] = 0;
  if (t == 0) S.errinv = 0;
}

// ------------------------------------------------------------------------------------------------
// phase 1: counts over the hot bytes of [b0, b1)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void p1_error_byte(const UstParams& P, Shared& S, unsigned b, long long i) {
  const unsigned code = b & 15u;
  if (!(b & UST_HOT_REVISION_HASH_ERROR) || !P.active) return;
  if (!(code == UST_STATE_UNKNOWN || code == UST_STATE_DONE || code == UST_STATE_POD_RESTART_REQUIRED || code == UST_STATE_FAILED)) return;
  if (__ldg(P.flags + i) & UST_F_POD_ORPHANED) return;  // orphaned pods never reach the hash lookup (common_manager.go:301-303)
  const unsigned long long key = UST_KEY(pass_of_state(code), (unsigned long long)i + 1ull);
  atomicMax(&S.errinv, ~key);
}

__device__ __forceinline__ void p1_word(const UstParams& P, Shared& S, uint32_t x, long long i, uint4& acc) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint4 inc = S.hotlut[(x >> (8 * k)) & 0xFFu];
    acc.x += inc.x; acc.y += inc.y; acc.z += inc.z; acc.w += inc.w;
  }
  if (x & 0x80808080u) {
#pragma unroll
    for (int k = 0; k < 4; k++) p1_error_byte(P, S, (x >> (8 * k)) & 0xFFu, i + k);
  }
}

__device__ __forceinline__ void p1_spill_thread(Shared& S, uint4& acc) {
  const uint32_t w[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int f = 0; f < 16; f++) {
    const unsigned v = (w[f >> 2] >> (8 * (f & 3))) & 0xFFu;
    if (v) atomicAdd(&S.cnt[f], v);
  }
  acc = make_uint4(0, 0, 0, 0);
}

__device__ void phase1(const UstParams& P, Shared& S, long long b0, long long b1) {
  const int t = threadIdx.x;
  uint4 acc = make_uint4(0, 0, 0, 0);
  int pending = 0;
  for (long long i = b0 + 16LL * t; i < b1; i += 16LL * kThreads) {
    if (i + 16 <= b1) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(P.hot + i));
      p1_word(P, S, h.x, i, acc);
      p1_word(P, S, h.y, i + 4, acc);
      p1_word(P, S, h.z, i + 8, acc);
      p1_word(P, S, h.w, i + 12, acc);
    } else {
      for (long long j = i; j < b1; j++) {  // ragged end of the array
        const unsigned b = P.hot[j];
        const uint4 inc = S.hotlut[b];
        acc.x += inc.x; acc.y += inc.y; acc.z += inc.z; acc.w += inc.w;
        p1_error_byte(P, S, b, j);
      }
    }
    pending += 16;
    if (pending > 255 - 16) { p1_spill_thread(S, acc); pending = 0; }  // byte lanes hold at most 255
  }
  // all threads converged: one REDUX per counter per warp, one shared atomic per warp
  {
    const uint32_t w[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int f = 0; f < 16; f++) {
      unsigned v = (w[f >> 2] >> (8 * (f & 3))) & 0xFFu;
      v = __reduce_add_sync(0xFFFFFFFFu, v);
      if ((t & 31) == 0 && v) atomicAdd(&S.cnt[f], v);
    }
  }
  __syncthreads();
  UstWorkspace* ws = P.ws;
  if (t < 14) {
    if (S.cnt[t]) atomicAdd(&ws->acc[t], (unsigned long long)S.cnt[t]);
  } else if (t == 14) {
    unsigned long long in = 0;
    for (int f = 0; f < 14; f++) in += S.cnt[f];
    const unsigned long long excluded = (unsigned long long)(b1 - b0) - in;
    if (excluded) atomicAdd(&ws->acc[UST_STATE_EXCLUDED], excluded);
  } else if (t == 15) {
    if (S.cnt[14]) atomicAdd(&ws->acc[UST_V_UNAVAILABLE], (unsigned long long)S.cnt[14]);
    if (S.cnt[15]) atomicAdd(&ws->acc[UST_V_CANDIDATES], (unsigned long long)S.cnt[15]);
  } else if (t == 32) {
    if (S.errinv) atomicMax(&ws->errinv, S.errinv);
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
  *P.out = c;
}

// ------------------------------------------------------------------------------------------------
// phase 2: per-node transition
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int load_dsrev(const UstParams& P, const Shared& S, uint32_t di) {
  if (P.n_ds <= UST_DS_SMEM_MAX) return S.dsrev[min(di, (uint32_t)P.n_ds)];
  return di < (uint32_t)P.n_ds ? __ldg(P.ds_rev + di) : 0;
}

// table entry for one node. hb = hot byte, extra = derived bits (slot grant, pod-list summaries)
__device__ __forceinline__ uint32_t node_entry(const UstParams& P, const Shared& S, uint32_t hb, uint32_t fl, int rev,
                                               uint32_t di, uint32_t extra) {
  const int dsr = load_dsrev(P, S, di);
  const bool synced = (di < (uint32_t)P.n_ds) && (rev == dsr);  // podRevisionHash == daemonsetRevisionHash (common_manager.go:318)
  uint32_t w = (fl & UST_F_INPUT_MASK) | ((hb >> 3) & (UST_W_SKIP | UST_W_UNSCHEDULABLE)) | extra;
  if (synced) w |= UST_W_SYNCED;
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

// fast path: full steps, CTA on one side of the slot budget, no abort, no pod lists
__device__ __forceinline__ void fast_tile(const UstParams& P, const Shared& S, long long base, uint32_t grant) {
  const int t = threadIdx.x;
  uint32_t h[kUnroll];
  uint4 f[kUnroll], r[kUnroll], d[kUnroll];
#pragma unroll
  for (int j = 0; j < kUnroll; j++) {
    const long long i = base + (long long)j * kStep + 4 * t;
    h[j] = ld_keep_u32(P.hot + i);
    f[j] = ld_stream_u4(P.flags + i);
    r[j] = ld_stream_u4(P.pod_rev + i);
    d[j] = ld_stream_u4(P.ds_idx + i);
  }
#pragma unroll
  for (int j = 0; j < kUnroll; j++) {
    const long long i = base + (long long)j * kStep + 4 * t;
    uint32_t e[4];
    e[0] = node_entry(P, S, h[j] & 0xFFu, f[j].x, (int)r[j].x, d[j].x, grant);
    e[1] = node_entry(P, S, (h[j] >> 8) & 0xFFu, f[j].y, (int)r[j].y, d[j].y, grant);
    e[2] = node_entry(P, S, (h[j] >> 16) & 0xFFu, f[j].z, (int)r[j].z, d[j].z, grant);
    e[3] = node_entry(P, S, h[j] >> 24, f[j].w, (int)r[j].w, d[j].w, grant);
    uint32_t next4, out4;
    uint2 act4;
    pack4(e, next4, act4, out4);
    __stcs(reinterpret_cast<uint32_t*>(P.next + i), next4);
    __stcs(reinterpret_cast<uint2*>(P.actions + i), act4);
    if (P.outcome) __stcs(reinterpret_cast<uint32_t*>(P.outcome + i), out4);
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
      const unsigned v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
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
  uint32_t e[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint32_t extra = gbits[k], f = fl[k];
    if (P.eval_pods && k < nvalid) {
      uint32_t clear_mask;
      extra |= pod_summary(P, hb[k], i0 + k, clear_mask);
      f &= ~clear_mask;
    }
    e[k] = node_entry(P, S, hb[k], f, rev[k], di[k], extra);
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

__device__ void phase2(const UstParams& P, Shared& S, long long b0, long long b1, unsigned chunk_cand) {
  // where does this chunk sit relative to the slot budget?
  const long long lo = S.cand_prefix, hi = S.cand_prefix + chunk_cand;
  const bool slotted = P.active && !P.requestor;
  const bool exact = slotted && chunk_cand != 0 && lo < S.budget && hi > S.budget;
  const uint32_t grant = (slotted && chunk_cand != 0 && hi <= S.budget) ? UST_W_GRANTED : 0u;
  const bool general = exact || S.abort_key != ~0ull || P.eval_pods;
  long long base = b0;
  if (!general) {
    for (; base + (long long)kStep * kUnroll <= b1; base += (long long)kStep * kUnroll) fast_tile(P, S, base, grant);
    long long running = 0;
    for (; base < b1; base += kStep) general_step<false>(P, S, base, b1, grant, running);
  } else if (!exact) {
    long long running = 0;
    for (; base < b1; base += kStep) general_step<false>(P, S, base, b1, grant, running);
  } else {
    long long running = 0;
    for (; base < b1; base += kStep) general_step<true>(P, S, base, b1, 0u, running);
  }
}

__device__ long long block_sum_cand_before(const UstParams& P, Shared& S, int chunk) {
  // exclusive prefix of per-chunk candidate counts (chunk order == slice order)
  const int t = threadIdx.x;
  long long s = 0;
  for (int c = t; c < chunk; c += kThreads) s += __ldcg(&P.ws->cand_cta[c]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  __shared__ long long part[kWarps];
  if ((t & 31) == 0) part[t >> 5] = s;
  __syncthreads();
  long long tot = 0;
  for (int w = 0; w < kWarps; w++) tot += part[w];
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

__device__ void finish(const UstParams& P, Shared& S, int chunks, bool reset_ws) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(&P.ws->depart, 1u);
    if (prev == (unsigned)chunks - 1u) {  // last CTA out: publish counters, restore the workspace invariant
      write_counters(P, S);
      if (reset_ws) {
        for (int i = 0; i < 18; i++) P.ws->acc[i] = 0;
        P.ws->errinv = 0;
      }
      P.ws->arrive = 0;
      P.ws->depart = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2) ust_fused_kernel(const __grid_constant__ UstParams P) {
  __shared__ Shared S;
  const int chunk = blockIdx.x, chunks = gridDim.x;
  const long long b0 = chunk_bound(P.n, chunk, chunks), b1 = chunk_bound(P.n, chunk + 1, chunks);
  stage_tables(P, S);
  __syncthreads();
  phase1(P, S, b0, b1);
  if (threadIdx.x == 0) P.ws->cand_cta[chunk] = S.cnt[15];
  // grid-wide barrier: every CTA is co-resident (cooperative launch)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&P.ws->arrive, 1u);
    while (ld_acquire_u32(&P.ws->arrive) < (unsigned)chunks) __nanosleep(40);
    __threadfence();
  }
  __syncthreads();
  load_local_vector(P, S);
  __syncthreads();
  if (threadIdx.x == 0) derive_scalars(P, S);
  const long long before = block_sum_cand_before(P, S, chunk);
  if (threadIdx.x == 0) S.cand_prefix += before;
  __syncthreads();
  phase2(P, S, b0, b1, S.cnt[15]);
  finish(P, S, chunks, true);
}

// split mode (multi-GPU with a host-launched collective between the phases, or pipelined uploads)
__global__ void __launch_bounds__(kThreads, 2) ust_phase1_kernel(const __grid_constant__ UstParams P) {
  __shared__ Shared S;
  const int chunk = blockIdx.x, chunks = gridDim.x;
  const long long b0 = chunk_bound(P.n, chunk, chunks), b1 = chunk_bound(P.n, chunk + 1, chunks);
  stage_tables(P, S);
  __syncthreads();
  phase1(P, S, b0, b1);
  if (threadIdx.x == 0) P.ws->cand_cta[chunk] = S.cnt[15];
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(&P.ws->depart, 1u) == (unsigned)chunks - 1u;
  }
  __syncthreads();
  if (last) {  // publish this shard's lanes of the exchange vector, restore the workspace invariant
    __threadfence();
    load_local_vector(P, S);
    __syncthreads();
    if (threadIdx.x < UST_V_LEN) P.xchg[threadIdx.x] = S.V[threadIdx.x];
    __syncthreads();
    if (threadIdx.x < 18) P.ws->acc[threadIdx.x] = 0;
    if (threadIdx.x == 0) { P.ws->errinv = 0; P.ws->depart = 0; }
  }
}

__global__ void __launch_bounds__(kThreads, 2) ust_phase2_kernel(const __grid_constant__ UstParams P) {
  __shared__ Shared S;
  const int chunk = P.chunk_begin + blockIdx.x, chunks = P.grid_chunks;
  const long long b0 = chunk_bound(P.n, chunk, chunks), b1 = chunk_bound(P.n, chunk + 1, chunks);
  stage_tables(P, S);
  if (threadIdx.x < UST_V_LEN) S.V[threadIdx.x] = P.xchg[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) derive_scalars(P, S);
  const long long before = block_sum_cand_before(P, S, chunk);
  if (threadIdx.x == 0) S.cand_prefix += before;
  __syncthreads();
  phase2(P, S, b0, b1, __ldcg(&P.ws->cand_cta[chunk]));
  finish(P, S, (int)gridDim.x, false);
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
