// ust_lut.h — per-policy transition tables consumed by the sm_100a kernels (ust_kernels.cu).
//
// The kernel never branches on a node's state. For every node it forms a 32-bit predicate word
//   w = (flags & UST_F_INPUT_MASK) | skip/unschedulable from the hot byte | derived bits
// and looks the result up in a table indexed by (state code, 9-bit window of w). Which window a state
// reads is fixed by the bit layout in include/ust.h; WHAT each (state, window) maps to depends on the
// policy and the manager options, so the table is rebuilt whenever those change (8192 entries, built
// on the host in microseconds, cached in the handle, 32 KiB in shared memory per CTA).
//
// Entry layout:  bits 0-15 actions (UST_A_*), bits 16-23 next state, bits 24-31 actuator outcome
// (UST_OUTCOME_NONE = no actuator ran) — byte-aligned so the kernel packs four nodes with PRMT.
#pragma once
#include <stdint.h>

#include "../../include/ust.h"

#define UST_LUT_WINDOW_BITS 9
#define UST_LUT_WINDOW (1u << UST_LUT_WINDOW_BITS)
#define UST_LUT_ENTRIES (16u * UST_LUT_WINDOW)

// bits of w that the kernel derives (never taken from the caller's flags word)
#define UST_W_SKIP (1u << 2)
#define UST_W_UNSCHEDULABLE (1u << 3)
#define UST_W_GRANTED (1u << 4)     /* upgrade slot available for this candidate (upgrade_inplace.go:87-99) */
#define UST_W_SYNCED (1u << 9)      /* pod revision hash == DaemonSet revision hash (common_manager.go:318) */
#define UST_W_PD_HAS (1u << 22)     /* numPodsToDelete != 0                    pod_manager.go:184 */
#define UST_W_PD_MISMATCH (1u << 23) /* numPodsCanDelete != numPodsToDelete    pod_manager.go:194 */
#define UST_W_DRAIN_ERROR (1u << 24) /* drain helper reports an error pod      drain_manager.go:121-128 */

// first bit of the window each state's transition reads (always >= 2: the kernel shifts by sh-2 so the
// table index comes out pre-multiplied by 4)
static constexpr int ust_window_shift[16] = {
    /* 0 unknown            */ 3,   // UNSCHED, UPG_REQ, SAFE_LOAD, ORPHANED, SYNCED
    /* 1 upgrade-required   */ 2,   // SKIP, UNSCHED, GRANTED, UPG_REQ
    /* 2 cordon-required    */ 2,
    /* 3 wait-for-jobs      */ 16,  // WAIT_*
    /* 4 pod-deletion       */ 22,  // PD_HAS, PD_MISMATCH
    /* 5 drain-required     */ 24,  // DRAIN_ERROR
    /* 6 node-maintenance   */ 20,  // NM_PRESENT, NM_READY
    /* 7 post-maintenance   */ 2,
    /* 8 pod-restart        */ 7,   // SAFE_LOAD, ORPHANED, SYNCED, POD_READY, INITIAL, REQUESTOR, TERMINATING, FAILING
    /* 9 validation         */ 6,   // VALIDATION_DONE, SAFE_LOAD, INITIAL, REQUESTOR
    /* 10 uncordon          */ 13,  // REQUESTOR
    /* 11 upgrade-done      */ 3,
    /* 12 upgrade-failed    */ 8,   // ORPHANED, SYNCED, POD_READY, INITIAL
    /* 13 other, 14 excluded, 15 reserved */ 2, 2, 2};

// position of each state's Process* pass in ApplyState's call order (upgrade_state.go:205-274);
// -1 = the state is never processed
static const int ust_pass_of_state[16] = {0, 2, 3, 4, 5, 6, 7, -1, 8, 10, 11, 1, 9, -1, -1, -1};

static inline uint32_t ust_lut_pack(unsigned state, unsigned next, unsigned actions, unsigned outcome) {
  if (next != state) actions |= UST_A_SET_STATE;
  return (actions & 0xFFFFu) | ((next & 0xFFu) << 16) | ((outcome & 0xFFu) << 24);
}

// updateNodeToUncordonOrDoneState (common_manager.go:673-708)
static inline void ust_uncordon_or_done(uint32_t w, unsigned* next, unsigned* actions) {
  bool requestor = (w & UST_F_REQUESTOR_MODE) != 0;
  *next = UST_STATE_UNCORDON_REQUIRED;
  if ((w & UST_F_INITIAL_STATE_ANNO) && !requestor) *next = UST_STATE_DONE;
  if (*next == UST_STATE_DONE || requestor) *actions |= UST_A_CLEAR_INITIAL_STATE_ANNO;
}

// One node's transition as a function of its state code, predicate word and the policy.
static inline uint32_t ust_transition(unsigned s, uint32_t w, const ust_policy* p) {
  unsigned next = s, a = 0, outcome = UST_OUTCOME_NONE;
  const bool orphan = (w & UST_F_POD_ORPHANED) != 0;
  const bool synced = !orphan && (w & UST_W_SYNCED);
  switch (s) {
    case UST_STATE_UNKNOWN:
    case UST_STATE_DONE:  // ProcessDoneOrUnknownNodes  common_manager.go:229-291
      if ((!synced && !orphan) || (w & UST_F_SAFE_LOAD) || (w & UST_F_UPGRADE_REQUESTED)) {
        if (w & UST_W_UNSCHEDULABLE) a |= UST_A_SET_INITIAL_STATE_ANNO;
        next = UST_STATE_UPGRADE_REQUIRED;
      } else if (s == UST_STATE_UNKNOWN) {
        next = UST_STATE_DONE;
      }
      break;
    case UST_STATE_UPGRADE_REQUIRED:
      if (w & UST_F_UPGRADE_REQUESTED) a |= UST_A_CLEAR_UPGRADE_REQUESTED;
      if (w & UST_W_SKIP) break;
      if (p->use_maintenance_operator) {  // upgrade_requestor.go:277-319
        a |= UST_A_NM_CREATE_OR_DELETE | UST_A_REQUESTOR_ANNO_CHANGE;
        next = UST_STATE_NODE_MAINTENANCE_REQUIRED;
      } else if ((w & UST_W_GRANTED) || (w & UST_W_UNSCHEDULABLE)) {  // upgrade_inplace.go:87-101
        next = UST_STATE_CORDON_REQUIRED;
      }
      break;
    case UST_STATE_CORDON_REQUIRED:  // common_manager.go:361-380
      a |= UST_A_CORDON;
      next = UST_STATE_WAIT_FOR_JOBS_REQUIRED;
      break;
    case UST_STATE_WAIT_FOR_JOBS_REQUIRED:  // common_manager.go:384-419
      if (!p->wait_selector_set) {
        next = p->pod_deletion_enabled ? UST_STATE_POD_DELETION_REQUIRED : UST_STATE_DRAIN_REQUIRED;
      } else {
        a |= UST_A_SCHEDULE_WAIT_CHECK;
        if (p->evaluate_actuators) {  // pod_manager.go:256-317, :331-368
          outcome = UST_STATE_WAIT_FOR_JOBS_REQUIRED;
          if (w & UST_F_WAIT_PODS_RUNNING) {
            if (p->wait_timeout_nonzero) {
              if (!(w & UST_F_WAIT_START_ANNO)) a |= UST_A_SET_WAIT_START;
              else if (w & UST_F_WAIT_START_INVALID) {}
              else if (w & UST_F_WAIT_TIMED_OUT) { outcome = UST_STATE_POD_DELETION_REQUIRED; a |= UST_A_CLEAR_WAIT_START; }
            }
          } else {
            a |= UST_A_CLEAR_WAIT_START;
            outcome = UST_STATE_POD_DELETION_REQUIRED;
          }
        }
      }
      break;
    case UST_STATE_POD_DELETION_REQUIRED:  // common_manager.go:424-453
      if (!p->pod_deletion_enabled) {
        next = UST_STATE_DRAIN_REQUIRED;
      } else {
        a |= UST_A_SCHEDULE_POD_EVICTION;
        if (p->evaluate_actuators) {  // pod_manager.go:176-220, :393-403
          if (!(w & UST_W_PD_HAS)) outcome = UST_STATE_POD_RESTART_REQUIRED;
          else if (w & UST_W_PD_MISMATCH) outcome = p->drain_enabled ? UST_STATE_DRAIN_REQUIRED : UST_STATE_FAILED;
          else outcome = UST_STATE_POD_RESTART_REQUIRED;
        }
      }
      break;
    case UST_STATE_DRAIN_REQUIRED:  // common_manager.go:329-357
      if (!p->drain_enabled) {
        next = UST_STATE_POD_RESTART_REQUIRED;
      } else {
        a |= UST_A_SCHEDULE_DRAIN;
        if (p->evaluate_actuators)  // drain_manager.go:106-131
          outcome = (w & UST_W_DRAIN_ERROR) ? UST_STATE_FAILED : UST_STATE_POD_RESTART_REQUIRED;
      }
      break;
    case UST_STATE_NODE_MAINTENANCE_REQUIRED:  // upgrade_requestor.go:416-452; in-place mode never touches it
      if (p->use_maintenance_operator) {
        if (!(w & UST_F_NM_PRESENT)) next = UST_STATE_UPGRADE_REQUIRED;
        else if (w & UST_F_NM_READY) next = UST_STATE_POD_RESTART_REQUIRED;
      }
      break;
    case UST_STATE_POD_RESTART_REQUIRED:  // common_manager.go:457-524
      if (!synced || orphan) {
        if (!(w & UST_F_POD_TERMINATING)) a |= UST_A_RESTART_DRIVER_POD;
      } else {
        if (w & UST_F_SAFE_LOAD) a |= UST_A_UNBLOCK_SAFE_LOAD;
        if (w & UST_F_POD_READY) {
          if (!p->validation_enabled) ust_uncordon_or_done(w, &next, &a);
          else next = UST_STATE_VALIDATION_REQUIRED;
        } else if (w & UST_F_POD_FAILING) {
          next = UST_STATE_FAILED;
        }
      }
      break;
    case UST_STATE_FAILED:  // common_manager.go:528-570 (no requestor-mode check here)
      if (synced && (w & UST_F_POD_READY)) {
        if (w & UST_F_INITIAL_STATE_ANNO) { next = UST_STATE_DONE; a |= UST_A_CLEAR_INITIAL_STATE_ANNO; }
        else next = UST_STATE_UNCORDON_REQUIRED;
      }
      break;
    case UST_STATE_VALIDATION_REQUIRED:  // common_manager.go:573-604
      if (w & UST_F_SAFE_LOAD) a |= UST_A_UNBLOCK_SAFE_LOAD;
      if (w & UST_F_VALIDATION_DONE) ust_uncordon_or_done(w, &next, &a);
      break;
    case UST_STATE_UNCORDON_REQUIRED:  // upgrade_inplace.go:124-147, upgrade_requestor.go:454-488
      if (!(w & UST_F_REQUESTOR_MODE)) { a |= UST_A_UNCORDON; next = UST_STATE_DONE; }
      else if (p->use_maintenance_operator) { next = UST_STATE_DONE; a |= UST_A_REQUESTOR_ANNO_CHANGE | UST_A_NM_CREATE_OR_DELETE; }
      break;
    default: break;  // post-maintenance-required, other, excluded: never processed
  }
  return ust_lut_pack(s, next, a, outcome);
}

// lut[s * 512 + key] for key = (w >> ust_window_shift[s]) & 511
static inline void ust_build_lut(const ust_policy* p, uint32_t* lut) {
  for (unsigned s = 0; s < 16; s++) {
    const int sh = ust_window_shift[s];
    for (uint32_t key = 0; key < UST_LUT_WINDOW; key++) {
      uint32_t w = (sh + UST_LUT_WINDOW_BITS >= 32) ? (uint32_t)(((uint64_t)key << sh) & 0xFFFFFFFFull) : (key << sh);
      lut[s * UST_LUT_WINDOW + key] = ust_transition(s, w, p);
    }
  }
}

// Pod-list table: for one workload pod, which actuator conditions it raises (index = pod_flags & 0x7FF).
// kubectl drain filter chain, k8s.io/kubectl v0.35.1 pkg/drain/filters.go (see oracle for the restatement).
#define UST_PODLUT_ENTRIES 2048u
#define UST_PODSUM_WAIT_RUNNING 0x01u
#define UST_PODSUM_TO_DELETE 0x02u    /* matches the deletion filter */
#define UST_PODSUM_CANNOT_DELETE 0x04u /* ... but the base filter chain keeps it */
#define UST_PODSUM_DRAIN_ERROR 0x08u

static inline bool ust_pod_chain_keeps(unsigned pf, bool force, bool delete_emptydir, bool* is_error) {
  const unsigned phase = pf & UST_POD_PHASE_MASK;
  const bool finished = phase == UST_PHASE_SUCCEEDED || phase == UST_PHASE_FAILED;
  *is_error = false;
  if ((pf & UST_POD_HAS_CONTROLLER) && (pf & UST_POD_CONTROLLED_BY_DS) && !finished) {
    if (pf & UST_POD_DS_MISSING) { if (!force) { *is_error = true; return true; } }
    else return true;  // skipped with a warning (IgnoreAllDaemonSets)
  }
  if (pf & UST_POD_MIRROR) return true;
  if ((pf & UST_POD_HAS_EMPTYDIR) && !finished && !delete_emptydir) { *is_error = true; return true; }
  if (!finished && !(pf & UST_POD_HAS_CONTROLLER) && !force) { *is_error = true; return true; }
  return false;
}

static inline void ust_build_pod_lut(const ust_policy* p, uint8_t* podlut) {
  for (unsigned pf = 0; pf < UST_PODLUT_ENTRIES; pf++) {
    unsigned r = 0;
    const unsigned phase = pf & UST_POD_PHASE_MASK;
    if ((pf & UST_POD_MATCH_WAIT_SELECTOR) && (phase == UST_PHASE_RUNNING || phase == UST_PHASE_PENDING)) r |= UST_PODSUM_WAIT_RUNNING;
    bool err;
    if (pf & UST_POD_MATCH_DELETION_FILTER) {
      r |= UST_PODSUM_TO_DELETE;
      if (ust_pod_chain_keeps(pf, p->pod_deletion_force != 0, p->pod_deletion_delete_emptydir != 0, &err)) r |= UST_PODSUM_CANNOT_DELETE;
    }
    if (pf & UST_POD_MATCH_DRAIN_SELECTOR) {
      ust_pod_chain_keeps(pf, p->drain_force != 0, p->drain_delete_emptydir != 0, &err);
      if (err) r |= UST_PODSUM_DRAIN_ERROR;
    }
    podlut[pf] = (uint8_t)r;
  }
}
