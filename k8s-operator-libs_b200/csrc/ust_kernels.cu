// ust_kernels.cu — the verification kernel of ApplyState and the auxiliary kernels of libust.so (sm_100a).
//
// ust_verify_kernel runs behind ust_stream_kernel (ust_stream.cu) on the same stream, launched with programmatic
// dependent launch: one small CTA per SM, resident (asleep in griddepcontrol.wait, 4.4 KiB transition table staged)
// while the streaming kernel runs. When that ends every CTA reads the shard's counters - on several GPUs: exchanges
// them through NVLink mailboxes, or takes the result of a host-launched NCCL all-reduce - derives the slot budget
// (GetUpgradesAvailable, common_manager.go:748-776) and judges the speculation the streaming kernel made. CTA 0 writes
// ust_counters. In the common case the speculation held, every output is final, and the kernel returns at once.
// Otherwise the CTAs re-evaluate, exactly, the tiles that need it: tiles before the cut with every upgrade candidate
// granted, tiles behind it with none, the cut tile with the ordered allocation of upgrade_inplace.go:71-109
// (candidate rank in slice order < slots left: warp-shuffle scan + per-step totals), and - when the call aborts -
// every tile with the reference's abort semantics (nodes the sequential passes had not reached stay untouched,
// common_manager.go:462-523).
// Around it: ust_pod_summary_kernel (pod lists -> one byte per node), ust_build_state*_kernel (BuildState),
// ust_patch_kernel / ust_feedback_kernel (delta updates, rollout simulation), ust_widen_kernel (packed host format).
#include <climits>

#include "ust_common.cuh"

using namespace ustd;

namespace {

constexpr int kThreads = UST_THREADS;
constexpr int kWarps = kThreads / 32;
// the verification kernel: 12 warps per CTA, one CTA per SM beside the resident streaming kernel (register bound below)
constexpr int kVThreads = UST_VERIFY_THREADS;
constexpr int kVWarps = kVThreads / 32;
constexpr int kVStep = kVThreads * 4;
constexpr uint32_t kLutBytes = UST_LUT_WORDS * sizeof(uint32_t);  // table + 16 {x, y} meta pairs
constexpr int kSimNone = INT32_MIN;        // rollout simulation: no start-time annotation
constexpr int kSimLongAgo = -(1 << 30);    // ... one that timed out before the simulation began
constexpr int kDsSmem = 64;  // the kernel sits next to the streaming kernel on every SM: keep its shared memory small

struct __align__(128) Shared {
  uint32_t lut[UST_LUT_ENTRIES];  // + meta directly behind it: filled by ONE bulk (TMA) copy
  uint2 meta[16];
  unsigned long long mbar;        // mbarrier the bulk copy completes on
  unsigned long long xflag;       // fused exchange: the flag word CTA 0 published (other CTAs)
  int dsrev[kDsSmem + 1];         // DaemonSet revisions (larger tables are read from global memory: this is the rare path)
  unsigned int warp_tot[kVWarps];
  // the verdict, CTA-uniform
  unsigned long long abort_key;   // ~0 = none
  long long node_offset;          // global index of this shard's node 0
  long long slots;                // ordered path: slots left at the start of the tile being evaluated
  int redo, cut, lo, hi;
  DecideShared D;
};

// byte lanes: B[0] = fields 0,2,4,6  B[1] = fields 1,3,5,7  B[2] = fields 8,10,12,14  B[3] = fields 9,11,13,15
__device__ __forceinline__ unsigned p1_field(const uint32_t (&B)[4], int f) {
  return (B[(f >> 3) * 2 + (f & 1)] >> (8 * ((f & 7) >> 1))) & 0xFFu;
}
__device__ __forceinline__ void widen(uint32_t& lo, uint32_t& hi, uint32_t (&B)[4]) {
  B[0] += lo & 0x0F0F0F0Fu;
  B[1] += (lo >> 4) & 0x0F0F0F0Fu;
  B[2] += hi & 0x0F0F0F0Fu;
  B[3] += (hi >> 4) & 0x0F0F0F0Fu;
  lo = hi = 0;
}

// ------------------------------------------------------------------------------------------------
// per-node transition
// ------------------------------------------------------------------------------------------------
// table entry for one node. hb = hot byte, extra = derived bits (slot grant, pod-list summaries)
__device__ __forceinline__ uint32_t node_entry(const UstParams& P, const Shared& S, bool ds_smem, uint32_t hb, uint32_t fl, int rev,
                                               uint32_t di, uint32_t extra) {
  uint32_t w = (fl & UST_F_INPUT_MASK) | ((hb >> 3) & (UST_W_SKIP | UST_W_UNSCHEDULABLE)) | extra;
  // podRevisionHash == daemonsetRevisionHash (common_manager.go:318); a missing DaemonSet never matches
  bool synced;
  if (ds_smem) synced = (di < (uint32_t)P.n_ds) && (rev == S.dsrev[min(di, (uint32_t)P.n_ds)]);
  else synced = di < (uint32_t)P.n_ds && rev == __ldg(P.ds_rev + di);
  if (synced) w |= UST_W_SYNCED;
  const uint2 m = S.meta[hb & 15u];
  const uint32_t off = (__funnelshift_r(w, 0u, m.x) & (m.x >> 16)) | m.y;
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

// One step of kVStep nodes starting at `base`, bounds-checked against b1 (the end of the tile / the shard).
// EXACT: the ordered slot allocation - candidate = upgrade-required && !skip; rank = exclusive count of candidates in
// slice order from the start of the tile (`running` carries it from step to step); granted iff rank < S.slots
// (upgrade_inplace.go:71-109). Otherwise the grant is uniform. Abort masking and pod-list summaries as in the
// streaming pass.
template <bool EXACT>
__device__ void general_step(const UstParams& P, Shared& S, long long base, long long b1, uint32_t grant, long long& running) {
  const int t = threadIdx.x;
  const long long i0 = base + 4 * t;
  const bool aborting = S.abort_key != ~0ull;
  uint32_t hb[4], fl[4], di[4];
  int rev[4];
  int nvalid = 0;
  if (i0 + 4 <= b1) {
    nvalid = 4;
    const uint32_t h = __ldg(reinterpret_cast<const uint32_t*>(P.hot + i0));
    const uint4 f = __ldcs(reinterpret_cast<const uint4*>(P.flags + i0)), r = __ldcs(reinterpret_cast<const uint4*>(P.pod_rev + i0)),
                d = __ldcs(reinterpret_cast<const uint4*>(P.ds_idx + i0));
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
    for (int w = 0; w < kVWarps; w++) {
      const unsigned v = S.warp_tot[w];
      if (w < (t >> 5)) before += v;
      step_total += v;
    }
    __syncthreads();
    long long rank = running + before + (incl - tc);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      gbits[k] = (c[k] && rank < S.slots) ? UST_W_GRANTED : 0u;
      rank += c[k];
    }
    running += step_total;
  }

  if (nvalid == 0) return;
  const bool ds_smem = P.n_ds <= kDsSmem;
  uint32_t e[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint32_t extra = gbits[k], f = fl[k];
    if (P.podsum && k < nvalid) {  // pod-list summary of the node (ust_pod_summary_kernel), see pods_apply()
      const uint32_t ps = P.podsum[i0 + k];
      f &= ~((ps & 0x10u) << 12);
      extra |= ((ps & 1u) << 16) | ((ps & 0xEu) << 21);
    }
    e[k] = node_entry(P, S, ds_smem, hb[k], f, rev[k], di[k], extra);
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

// A span of full steps with a uniform grant and no abort, software-pipelined: while the two steps (2048 nodes) of one
// iteration are evaluated, the loads of the next two are already in flight (two register buffers, the loop is unrolled
// over them). b0, b1: multiples of kVStep apart (the caller peels the ragged end).
struct SpanTile {
  uint32_t h[2], ps[2];
  uint4 f[2], r[2], d[2];
};
__device__ __forceinline__ void span_load(const UstParams& P, SpanTile& T, long long base, long long b1) {
  const int t = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const long long i = base + (long long)j * kVStep + 4 * t;
    const bool v = i < b1;   // warp-uniform: spans are multiples of kVStep
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    T.h[j] = v ? __ldg(reinterpret_cast<const uint32_t*>(P.hot + i)) : 0x0E0E0E0Eu;
    T.ps[j] = (v && P.podsum) ? __ldcs(reinterpret_cast<const uint32_t*>(P.podsum + i)) : 0u;
    T.f[j] = v ? __ldcs(reinterpret_cast<const uint4*>(P.flags + i)) : zero;
    T.r[j] = v ? __ldcs(reinterpret_cast<const uint4*>(P.pod_rev + i)) : zero;
    T.d[j] = v ? __ldcs(reinterpret_cast<const uint4*>(P.ds_idx + i)) : zero;
  }
}
__device__ __forceinline__ void span_eval(const UstParams& P, Shared& S, const SpanTile& T, long long base, long long b1, uint32_t grant,
                                          bool ds_smem) {
  const int t = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const long long i = base + (long long)j * kVStep + 4 * t;
    if (i >= b1) continue;
    const uint32_t fl[4] = {T.f[j].x, T.f[j].y, T.f[j].z, T.f[j].w};
    const uint32_t rv[4] = {T.r[j].x, T.r[j].y, T.r[j].z, T.r[j].w};
    const uint32_t dv[4] = {T.d[j].x, T.d[j].y, T.d[j].z, T.d[j].w};
    uint32_t e[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t hb = (T.h[j] >> (8 * k)) & 0xFFu, p = (T.ps[j] >> (8 * k)) & 0xFFu;
      const uint32_t fk = fl[k] & ~((p & 0x10u) << 12);
      e[k] = node_entry(P, S, ds_smem, hb, fk, (int)rv[k], dv[k], grant | ((p & 1u) << 16) | ((p & 0xEu) << 21));
    }
    uint32_t next4, out4;
    uint2 act4;
    pack4(e, next4, act4, out4);
    __stcs(reinterpret_cast<uint32_t*>(P.next + i), next4);
    __stcs(reinterpret_cast<uint2*>(P.actions + i), act4);
    if (P.outcome) __stcs(reinterpret_cast<uint32_t*>(P.outcome + i), out4);
  }
}
__device__ void uniform_span(const UstParams& P, Shared& S, long long b0, long long b1, uint32_t grant) {
  const bool ds_smem = P.n_ds <= kDsSmem;
  constexpr long long kIter = 2LL * kVStep;
  SpanTile A, B;
  span_load(P, A, b0, b1);
  for (long long base = b0; base < b1; base += 2 * kIter) {
    span_load(P, B, base + kIter, b1);           // (all-invalid past the end: no loads are issued)
    span_eval(P, S, A, base, b1, grant, ds_smem);
    span_load(P, A, base + 2 * kIter, b1);
    span_eval(P, S, B, base + kIter, b1, grant, ds_smem);
  }
}

// candidates (upgrade-required && !skip, upgrade_inplace.go:82) among the nodes [b0, b1) - hot bytes only; b0 and
// b1 are multiples of kVStep apart inside one tile. Whole CTA; every thread returns the total.
__device__ long long count_candidates(const UstParams& P, Shared& S, long long b0, long long b1) {
  const int t = threadIdx.x;
  unsigned c = 0;
  for (long long i = b0 + 16LL * t; i < b1; i += 16LL * kVThreads) {  // b0: multiple of 1024, 16 t < 4096: aligned 16-byte loads
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(P.hot + i));
    c += __popc(cand_mask4(x.x)) + __popc(cand_mask4(x.y)) + __popc(cand_mask4(x.z)) + __popc(cand_mask4(x.w));
  }
  c = __reduce_add_sync(kFull, c);
  __syncthreads();
  if ((t & 31) == 0) S.warp_tot[t >> 5] = c;
  __syncthreads();
  long long tot = 0;
#pragma unroll
  for (int w = 0; w < kVWarps; w++) tot += S.warp_tot[w];
  __syncthreads();
  return tot;
}

// Re-evaluate the steps [s0, s1) (kVStep nodes each) of one tile exactly, given where the slot budget cuts.
__device__ void redo_steps(const UstParams& P, Shared& S, int tile, int s0, int s1) {
  const long long t0 = (long long)tile * P.tile_nodes;
  long long t1 = t0 + P.tile_nodes;
  if (t1 > P.n) t1 = P.n;
  const long long b0 = t0 + (long long)s0 * kVStep;
  long long b1 = t0 + (long long)s1 * kVStep;
  if (b1 > t1) b1 = t1;
  if (b0 >= b1) return;
  const bool slotted = P.active && !P.requestor;
  const bool aborting = S.abort_key != ~0ull;
  if (slotted && tile == S.cut) {  // the cut tile: S.slots of its candidates get a slot, in slice order
    long long running = s0 > 0 ? count_candidates(P, S, t0, b0) : 0;  // candidates of the tile before this piece
    for (long long base = b0; base < b1; base += kVStep) general_step<true>(P, S, base, b1, 0u, running);
    return;
  }
  const uint32_t grant = (slotted && tile < S.cut) ? UST_W_GRANTED : 0u;
  long long running = 0, full_end = b0;
  if (!aborting) {
    full_end = b0 + ((b1 - b0) / kVStep) * kVStep;
    if (full_end > b0) uniform_span(P, S, b0, full_end, grant);
  }
  for (long long base = full_end; base < b1; base += kVStep) general_step<false>(P, S, base, b1, grant, running);
}

// Re-evaluate the tiles [ta, tb) - a contiguous run owned by one CTA when many tiles are redone. Without an abort and
// with tiles that are whole steps, the tiles before the cut and the tiles behind it are two node ranges that go
// through the pipelined span as a whole; the cut tile takes the ordered path.
__device__ void redo_range(const UstParams& P, Shared& S, int ta, int tb) {
  const int tn = P.tile_nodes;
  const int steps_per_tile = (tn + kVStep - 1) / kVStep;
  const bool aborting = S.abort_key != ~0ull;
  if (aborting || tn % kVStep != 0) {
    for (int tile = ta; tile < tb; tile++) { redo_steps(P, S, tile, 0, steps_per_tile); __syncthreads(); }
    return;
  }
  const bool slotted = P.active && !P.requestor;
  const int cut = slotted ? S.cut : 0x7FFFFFFF;
  for (int part = 0; part < 2; part++) {
    const int x = part == 0 ? ta : (cut + 1 > ta ? cut + 1 : ta);
    const int y = part == 0 ? (cut < tb ? cut : tb) : tb;
    if (x >= y) continue;
    const uint32_t grant = (slotted && part == 0) ? UST_W_GRANTED : 0u;
    const long long b0 = (long long)x * tn;
    long long b1 = (long long)y * tn;
    if (b1 > P.n) b1 = P.n;
    const long long full_end = b0 + ((b1 - b0) / kVStep) * kVStep;
    if (full_end > b0) uniform_span(P, S, b0, full_end, grant);
    long long running = 0;
    for (long long base = full_end; base < b1; base += kVStep) general_step<false>(P, S, base, b1, grant, running);
  }
  if (cut >= ta && cut < tb) { __syncthreads(); redo_steps(P, S, cut, 0, steps_per_tile); }
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// Registers: an SM sub-partition has 16384. The streaming kernel's CTA puts 4 of its 13 warps (64 registers a thread) on
// one sub-partition = 8192; this kernel's 12 warps come 3 to a sub-partition, so they must stay within 8192 / 96 = 85
// registers a thread to be resident BESIDE the streaming CTA. At 96 this kernel only got onto an SM when the streaming
// CTA left it - and the next call's streaming kernel, whose launch waits for every CTA here to have started, with it.
__global__ void __maxnreg__(80) ust_verify_kernel(const __grid_constant__ UstParams P) {
  __shared__ Shared S;
  const int t = threadIdx.x;
  // prologue (overlaps the streaming kernel): the transition table (4.4 KiB) by one TMA bulk copy - it was uploaded by
  // a copy, not by a kernel, so it need not wait
  if (t == 0) {
    mbar_init(&S.mbar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&S.mbar, kLutBytes);
    bulk_g2s(S.lut, P.lut, kLutBytes, &S.mbar);
  }
  if (P.stamps && t == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) P.ws->dbg2[blockIdx.x == 0 ? 8 : 9] = now_ns();
  griddep_launch_dependents();  // the next call's streaming kernel may become resident (it waits for this grid itself)
  griddep_wait();               // the streaming kernel (and, split mode, the collective) has completed
  const bool lead = blockIdx.x == 0;
  if (P.stamps && lead && t == 0) P.ws->dbg2[0] = now_ns();
  bool comm_ok = true;
  if (P.split) {
    if (t < UST_V_LEN) S.D.V[t] = P.xchg[t];
  } else {
    load_local_vector(P, S.D);
    if (t < 32) {
      __syncwarp();
      if (t == 14) fix_excluded_lane(P, S.D);
    }
    if (P.fused_exchange) {
      // Only CTA 0 polls the peers' mailbox words. (Every CTA polling them - 148 x 84 x world threads spinning on a few
      // L2 lines that the peers' NVLink writes must get into - cost 17 us a step at 8 GPUs.) The others wait for one
      // flag, one thread each.
      __syncthreads();
      const int par = (int)(P.epoch & 1);
      const unsigned long long tag = (unsigned long long)(unsigned)P.epoch << 32;
      if (lead) {
        comm_ok = exchange_vector(P, S.D, true);
        // the sum and its flag stay on this GPU: gpu scope, written by warps that issued no NVLink store (a system-scope
        // release - or a fence by a thread with peer stores in flight - waits for the peers' acknowledgements)
        if (t < UST_V_LEN) P.ws->xsum[par][t] = S.D.V[t];
        __syncthreads();
        if (t == 0) st_release_gpu(reinterpret_cast<long long*>(&P.ws->xflag[par]), (long long)(tag | (comm_ok ? 1ull : 0ull)));
      } else {
        if (t == 0) {
          const unsigned long long t0 = now_ns();
          unsigned long long v = (unsigned long long)ld_acquire_gpu(reinterpret_cast<const long long*>(&P.ws->xflag[par]));
          while ((v >> 32) != (tag >> 32)) {
            if (now_ns() - t0 > 2 * kCommTimeoutNs) { v = tag; break; }
            __nanosleep(100);
            v = (unsigned long long)ld_acquire_gpu(reinterpret_cast<const long long*>(&P.ws->xflag[par]));
          }
          S.xflag = v;
        }
        __syncthreads();
        comm_ok = (S.xflag & 1ull) != 0;
        if (t < UST_V_LEN) S.D.V[t] = __ldcg(&P.ws->xsum[par][t]);
      }
    }
  }
  if (t == 0) S.D.spec_cut = __ldcg(&P.ws->spec_used[P.parity]);
  __syncthreads();
  if (P.stamps && lead && t == 0) P.ws->dbg2[1] = now_ns();
  decide(P, S.D, lead, !comm_ok);
  if (P.stamps && lead && t == 0) P.ws->dbg2[2] = now_ns();
  const int redo = S.D.redo;
  mbar_wait(&S.mbar, 0);  // never leave with the bulk copy in flight (it landed long ago)
  if (redo == 0) return;  // the speculation held: every output of the streaming kernel is final
  if (t == 0) {
    S.redo = redo; S.cut = S.D.cut; S.lo = S.D.lo; S.hi = S.D.hi; S.slots = S.D.slots_left;
    S.abort_key = S.D.abort_key; S.node_offset = S.D.node_offset;
  }
  if (P.n_ds <= kDsSmem)
    for (int i = t; i <= P.n_ds; i += kVThreads) S.dsrev[i] = i < P.n_ds ? __ldg(P.ds_rev + i) : 0;
  __syncthreads();
  int first = S.lo, last = S.hi;
  if (redo == 2) { first = 0; last = P.n_tiles - 1; }
  const int m = last - first + 1;
  const int steps_per_tile = (P.tile_nodes + kVStep - 1) / kVStep;
  if (m >= (int)gridDim.x) {
    // many tiles: a contiguous run per CTA
    const int per = (m + (int)gridDim.x - 1) / (int)gridDim.x;
    const int ta = first + (int)blockIdx.x * per, tb = ta + per < last + 1 ? ta + per : last + 1;
    if (ta < tb) redo_range(P, S, ta, tb);
  } else {
    // a few tiles (the steady state: the one tile the budget cuts through): one step per CTA, so that the redo costs
    // one load round trip instead of one per step
    const int items = m * steps_per_tile;
    for (int it = (int)blockIdx.x; it < items; it += (int)gridDim.x) {
      const int tile = first + it / steps_per_tile, st = it % steps_per_tile;
      redo_steps(P, S, tile, st, st + 1);
      __syncthreads();
    }
  }
  if (P.stamps && lead && t == 0) P.ws->dbg2[3] = now_ns();
}

// Pod-list summaries (rows 12-14 of the scope table: pod_manager.go:256-391, :122-229, drain_manager.go:58-139).
// Only nodes whose actuator would look at its pods have their list read: wait-for-jobs, pod-deletion and
// drain-required nodes - everything else costs the hot byte. A CTA takes blocks of kPodBlock consecutive nodes:
// it compacts the nodes that need their list (in node order) into shared memory, then every thread walks the list
// of one such node with aligned 16-byte loads (8 pods each, all loads of a pass in flight together). A pod counts
// only if it matches the selector of the node's own actuator (wait-for-completion selector, deletion filter or drain
// selector: one bit of pod_flags each); those are mapped through the per-policy table of the pod's own eight bits
// (256 bytes in shared memory) and OR-ed, and the list is dropped as soon as every bit the node's state reads is set
// (round 2: 82 -> 62 us; the lookups of the 2 KiB table had the LSU data pipe at 76 % of peak). One thread per list
// rather than one warp (round-1 measurement: the warp-per-node formulation executed 15x the instructions).
// Neighbouring threads own neighbouring lists, so their loads share sectors. Output: one byte per node for the streaming pass (layout: pods_apply()), written
// coalesced per block.
constexpr int kPodBlock = 4096;            // nodes per block = 16 per thread
constexpr int kPodChunks = 6;              // 16-byte loads in flight per thread and pass (48 pods: a typical list in one pass)

__global__ void __launch_bounds__(kThreads, 6) ust_pod_summary_kernel(long long n, int active, const uint8_t* __restrict__ hot,
                                                                      const int32_t* __restrict__ pod_off,
                                                                      const uint16_t* __restrict__ pod_flags, long long n_pods,
                                                                      const uint8_t* __restrict__ podlut, uint8_t* __restrict__ podsum) {
  __shared__ __align__(16) uint8_t lut[256];   // what a pod raises, by its own eight bits (ust_lut.h: ust_build_pod_lut256)
  __shared__ __align__(16) uint8_t res[kPodBlock];   // summary byte per node of the block (bits 6-7: state - 3 while in work)
  __shared__ unsigned short list[kPodBlock];         // block-local indices of the nodes whose list is read
  __shared__ int cnt;
  const int t = threadIdx.x, lane = t & 31;
  if (t < 64) reinterpret_cast<uint32_t*>(lut)[t] = __ldg(reinterpret_cast<const uint32_t*>(podlut) + t);
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
      // the node's actuator only asks about pods that match ITS selector: wait-for-completion (bit 9), the deletion
      // filter (bit 8) or the drain selector (bit 10) - the others are not even looked up
      const unsigned tag = res[li];
      const unsigned s = UST_STATE_WAIT_FOR_JOBS_REQUIRED + (tag >> 6);
      const uint32_t sel = s == UST_STATE_WAIT_FOR_JOBS_REQUIRED ? (uint32_t)UST_POD_MATCH_WAIT_SELECTOR
                         : (s == UST_STATE_POD_DELETION_REQUIRED ? (uint32_t)UST_POD_MATCH_DELETION_FILTER : (uint32_t)UST_POD_MATCH_DRAIN_SELECTOR);
      // ... and once every bit the node's state reads is set, the rest of the list cannot change the answer
      const unsigned sat = s == UST_STATE_WAIT_FOR_JOBS_REQUIRED ? UST_PODSUM_WAIT_RUNNING
                         : (s == UST_STATE_POD_DELETION_REQUIRED ? (UST_PODSUM_TO_DELETE | UST_PODSUM_CANNOT_DELETE) : UST_PODSUM_DRAIN_ERROR);
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
            if (cb + u < nchunks && (r & sat) != sat) {
              const uint32_t wv[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
              for (int e = 0; e < 8; e++) {
                const uint32_t f = (e & 1) ? wv[e >> 1] >> 16 : wv[e >> 1];
                if ((f & sel) && (unsigned)(rel + e) < (unsigned)len) r |= lut[f & 255u];
              }
            }
            rel += 8;
          }
        }
      } else {  // the list ends within the last 16 bytes of the whole array: plain 2-byte loads
        for (int p = p0; p < p1; p++) {
          const uint32_t f = __ldg(pod_flags + p);
          if (f & sel) r |= lut[f & 255u];
        }
      }
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
  // fields 0..13 per state code, 14 unavailable, 15 candidates -> ws->bs_acc[0..13], [16], [17]
  for (int o = 16; o > 0; o >>= 1) excluded += __shfl_xor_sync(kFull, excluded, o);
  if ((t & 31) == 0 && excluded) atomicAdd(&ws->bs_acc[UST_STATE_EXCLUDED], (unsigned long long)excluded);
  if (t < 14) { if (cnt[t]) atomicAdd(&ws->bs_acc[t], (unsigned long long)cnt[t]); }
  else if (t == 14) { if (cnt[14]) atomicAdd(&ws->bs_acc[16], (unsigned long long)cnt[14]); }
  else if (t == 15) { if (cnt[15]) atomicAdd(&ws->bs_acc[17], (unsigned long long)cnt[15]); }
  if (in_smem)
    for (int i = t; i < n_ds; i += kThreads)
      if (cnt_ds[i]) atomicAdd(&ds_count[i], (unsigned long long)cnt_ds[i]);
}

__global__ void ust_build_state_finish_kernel(int n_ds, const int32_t* ds_desired, unsigned long long* ds_count,
                                              UstWorkspace* ws, ust_counters* out) {
  ust_counters c;
  for (int i = 0; i < 16; i++) c.hist[i] = (long long)ws->bs_acc[i];
  c.unavailable = (long long)ws->bs_acc[16];
  c.candidates = (long long)ws->bs_acc[17];
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
  for (int i = 0; i < 18; i++) ws->bs_acc[i] = 0;
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

// Rollout simulation: the state feedback between two reconciles. Untimed (sp.timed == 0): "ideal actuators" - every call
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
                                                                const uint8_t* __restrict__ outcome, const ust_counters* step,
                                                                const UstSimParams sp, int32_t* entered, int32_t* wait_start,
                                                                int32_t* valid_start) {
  if (step->error_code != UST_OK) return;  // the reconcile returned an error: nothing it decided is fed back
  const long long stride = (long long)gridDim.x * kThreads;
  const long long now = sp.now, now_next = sp.now + sp.dt;
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
    // requestor mode (upgrade_requestor.go:277-319, :454-488): the annotation and the NodeMaintenance object
    if (a & UST_A_REQUESTOR_ANNO_CHANGE) f = s == UST_STATE_UPGRADE_REQUIRED ? (f | UST_F_REQUESTOR_MODE) : (f & ~UST_F_REQUESTOR_MODE);
    if (a & UST_A_NM_CREATE_OR_DELETE) {
      if (s == UST_STATE_UPGRADE_REQUIRED) {
        f |= UST_F_NM_PRESENT;
      } else {  // the object goes, and with it the maintenance operator's cordon
        f &= ~(UST_F_NM_PRESENT | UST_F_NM_READY);
        b &= ~UST_HOT_UNSCHEDULABLE;
      }
    }
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
    int ent = 0, ws = 0, vs = 0;
    if (sp.timed) {
      ent = entered[i]; ws = wait_start[i]; vs = valid_start[i];
      if (a & UST_A_SET_WAIT_START) ws = (int)now;                 // annotation = currentTime (pod_manager.go:339)
      if (a & UST_A_CLEAR_WAIT_START) ws = kSimNone;
      // Validate() of a node whose validation pod is not ready runs handleTimeout (validation_manager.go:139-175)
      if (s == UST_STATE_VALIDATION_REQUIRED && ns == UST_STATE_VALIDATION_REQUIRED && !(f & UST_F_VALIDATION_DONE)) {
        if (vs == kSimNone) vs = (int)now;
        else if (now > (long long)vs + sp.validation_timeout) { ns = UST_STATE_FAILED; vs = kSimNone; }
      }
      if (ns != UST_STATE_VALIDATION_REQUIRED) vs = kSimNone;      // the annotation is removed once the pod is ready (:104-110)
      if (ns != s) ent = (int)now;
    }
    if (ns == UST_STATE_WAIT_FOR_JOBS_REQUIRED) {
      if (!sp.timed) f &= ~UST_F_WAIT_PODS_RUNNING;
      else {
        if (ns != s) f = sp.job_seconds > 0 ? (f | UST_F_WAIT_PODS_RUNNING) : (f & ~UST_F_WAIT_PODS_RUNNING);
        if (now_next >= (long long)ent + sp.job_seconds) f &= ~UST_F_WAIT_PODS_RUNNING;
        const bool timed_out = ws != kSimNone && now_next > (long long)ws + sp.wait_timeout;
        f = timed_out ? (f | UST_F_WAIT_TIMED_OUT) : (f & ~UST_F_WAIT_TIMED_OUT);
      }
    }
    if (ns == UST_STATE_POD_RESTART_REQUIRED) {
      f &= ~UST_F_POD_TERMINATING;
      if (!(f & UST_F_POD_FAILING)) f |= UST_F_POD_READY;
    }
    if (ns == UST_STATE_VALIDATION_REQUIRED) {
      if (!sp.timed) f |= UST_F_VALIDATION_DONE;
      else {
        const bool ready = sp.validation_seconds >= 0 && now_next >= (long long)ent + sp.validation_seconds;
        f = ready ? (f | UST_F_VALIDATION_DONE) : (f & ~UST_F_VALIDATION_DONE);
      }
    }
    if (ns == UST_STATE_NODE_MAINTENANCE_REQUIRED && (f & UST_F_NM_PRESENT)) {
      // the maintenance operator: cordon + drain, then Ready (maintenance_seconds after the object was created)
      const bool ready = !sp.timed || now_next >= (long long)ent + sp.maintenance_seconds;
      if (ready) { f |= UST_F_NM_READY; b |= UST_HOT_UNSCHEDULABLE; }
    }
    hot[i] = (uint8_t)((b & 0xF0u) | (ns & 15u));
    flags[i] = f;
    pod_rev[i] = rev;
    if (sp.timed) { entered[i] = ent; wait_start[i] = ws; valid_start[i] = vs; }
  }
}

// the per-node clocks of a timed simulation at time 0: every node "entered" its state then; a wait-start annotation that
// is already there started then, or - when the snapshot says it has timed out - long ago
__global__ void __launch_bounds__(kThreads) ust_sim_init_kernel(long long n, const uint32_t* __restrict__ flags, int32_t* entered,
                                                                int32_t* wait_start, int32_t* valid_start) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const uint32_t f = flags[i];
    entered[i] = 0;
    wait_start[i] = !(f & UST_F_WAIT_START_ANNO) ? kSimNone : ((f & UST_F_WAIT_TIMED_OUT) ? kSimLongAgo : 0);
    valid_start[i] = kSimNone;
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
// Sparse outputs of a delta call: the nodes whose (next_state, actions) differ from the previous call's,
// compacted in node order. Three launches: per-block counts, a one-CTA scan of the block counts, the ordered write.
// 6 B/node read twice; the alternative is 3 B/node over PCIe.
constexpr int kDiffBlock = 4096;  // nodes per CTA: 16 per thread
__device__ __forceinline__ unsigned diff_mask16(const uint8_t* next, const uint16_t* act, const uint8_t* pnext, const uint16_t* pact,
                                                long long i0, long long n) {
  unsigned m = 0;
  if (i0 + 16 <= n) {
    const uint4 a = __ldcs(reinterpret_cast<const uint4*>(next + i0)), b = __ldcs(reinterpret_cast<const uint4*>(pnext + i0));
    const uint4 c0 = __ldcs(reinterpret_cast<const uint4*>(act + i0)), c1 = __ldcs(reinterpret_cast<const uint4*>(act + i0 + 8));
    const uint4 d0 = __ldcs(reinterpret_cast<const uint4*>(pact + i0)), d1 = __ldcs(reinterpret_cast<const uint4*>(pact + i0 + 8));
    const uint32_t x[4] = {a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w};
    const uint32_t y[8] = {c0.x ^ d0.x, c0.y ^ d0.y, c0.z ^ d0.z, c0.w ^ d0.w, c1.x ^ d1.x, c1.y ^ d1.y, c1.z ^ d1.z, c1.w ^ d1.w};
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const bool dn = ((x[k >> 2] >> (8 * (k & 3))) & 0xFFu) != 0;
      const bool da = ((y[k >> 1] >> (16 * (k & 1))) & 0xFFFFu) != 0;
      m |= (dn || da ? 1u : 0u) << k;
    }
  } else {
    for (int k = 0; k < 16; k++)
      if (i0 + k < n && (next[i0 + k] != pnext[i0 + k] || act[i0 + k] != pact[i0 + k])) m |= 1u << k;
  }
  return m;
}
__global__ void __launch_bounds__(kThreads) ust_diff_count_kernel(long long n, const uint8_t* __restrict__ next, const uint16_t* __restrict__ act,
                                                                  const uint8_t* __restrict__ pnext, const uint16_t* __restrict__ pact,
                                                                  unsigned int* __restrict__ block_count) {
  __shared__ unsigned int tot;
  if (threadIdx.x == 0) tot = 0;
  __syncthreads();
  const long long i0 = (long long)blockIdx.x * kDiffBlock + 16 * threadIdx.x;
  unsigned c = i0 < n ? __popc(diff_mask16(next, act, pnext, pact, i0, n)) : 0u;
  c = __reduce_add_sync(kFull, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&tot, c);
  __syncthreads();
  if (threadIdx.x == 0) block_count[blockIdx.x] = tot;
}
// exclusive scan of the block counts in place (one CTA); total -> *n_out
__global__ void __launch_bounds__(1024) ust_diff_scan_kernel(int blocks, unsigned int* block_count, long long* n_out) {
  __shared__ unsigned long long part[32];
  __shared__ unsigned long long carry;
  const int t = threadIdx.x;
  if (t == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < blocks; b0 += 1024) {
    const int b = b0 + t;
    const unsigned long long v = b < blocks ? block_count[b] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(kFull, incl, o);
      if ((t & 31) >= o) incl += u;
    }
    if ((t & 31) == 31) part[t >> 5] = incl;
    __syncthreads();
    unsigned long long before = carry, tot = 0;
    for (int w = 0; w < 32; w++) { if (w < (t >> 5)) before += part[w]; tot += part[w]; }
    // offsets fit 32 bits per launch range: the API caps a sparse call at 2^31 changed outputs
    if (b < blocks) block_count[b] = (unsigned int)(before + incl - v);
    __syncthreads();
    if (t == 0) carry += tot;
    __syncthreads();
  }
  if (t == 0) *n_out = (long long)carry;
}
__global__ void __launch_bounds__(kThreads) ust_diff_write_kernel(long long n, const uint8_t* __restrict__ next, const uint16_t* __restrict__ act,
                                                                  const uint8_t* __restrict__ pnext, const uint16_t* __restrict__ pact,
                                                                  const unsigned int* __restrict__ block_off, long long cap,
                                                                  long long* __restrict__ out_idx, uint8_t* __restrict__ out_next,
                                                                  uint16_t* __restrict__ out_act) {
  __shared__ unsigned int wtot[kWarps];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const long long i0 = (long long)blockIdx.x * kDiffBlock + 16 * t;
  const unsigned m = i0 < n ? diff_mask16(next, act, pnext, pact, i0, n) : 0u;
  const unsigned c = __popc(m);
  unsigned incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned u = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) wtot[warp] = incl;
  __syncthreads();
  unsigned before = 0;
  for (int w = 0; w < warp; w++) before += wtot[w];
  long long pos = (long long)block_off[blockIdx.x] + before + incl - c;
  unsigned mm = m;
  while (mm) {
    const int k = __ffs(mm) - 1;
    mm &= mm - 1;
    if (pos < cap) {
      out_idx[pos] = i0 + k;
      out_next[pos] = next[i0 + k];
      out_act[pos] = act[i0 + k];
    }
    pos++;
  }
}

}  // namespace

int ust_launch_verify(const UstParams& p, int grid, void* stream, int pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kVThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, ust_verify_kernel, p);
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
                        const ust_counters* step, const UstSimParams& sp, int32_t* entered, int32_t* wait_start,
                        int32_t* valid_start, int grid, void* stream) {
  if (n <= 0) return 0;
  ust_feedback_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, hot, flags, pod_rev, ds_idx, n_ds, ds_rev, next, actions, outcome, step,
                                                                   sp, entered, wait_start, valid_start);
  return (int)cudaGetLastError();
}
int ust_launch_sim_init(long long n, const uint32_t* flags, int32_t* entered, int32_t* wait_start, int32_t* valid_start, int grid,
                        void* stream) {
  if (n <= 0) return 0;
  ust_sim_init_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, flags, entered, wait_start, valid_start);
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
int ust_launch_diff(long long n, const uint8_t* next, const uint16_t* actions, const uint8_t* prev_next, const uint16_t* prev_actions,
                    unsigned int* block_count, long long* n_out, long long cap, long long* out_idx, uint8_t* out_next,
                    uint16_t* out_actions, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long blocks = (n + kDiffBlock - 1) / kDiffBlock;
  if (blocks > 0) ust_diff_count_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(n, next, actions, prev_next, prev_actions, block_count);
  ust_diff_scan_kernel<<<1, 1024, 0, st>>>((int)blocks, block_count, n_out);
  if (blocks > 0)
    ust_diff_write_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(n, next, actions, prev_next, prev_actions, block_count, cap, out_idx, out_next, out_actions);
  return (int)cudaGetLastError();
}
int ust_diff_blocks(long long n) { return (int)((n + kDiffBlock - 1) / kDiffBlock); }
