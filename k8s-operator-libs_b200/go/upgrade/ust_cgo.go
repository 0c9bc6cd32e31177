// Package upgrade — cgo shim over libust.so for github.com/NVIDIA/k8s-operator-libs/pkg/upgrade.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain (DESIGN.md §1). The file shows the
// binding a maintainer adds next to pkg/upgrade/upgrade_state.go so that ClusterUpgradeStateManagerImpl keeps its
// interface (upgrade_state.go:35-53), its types (common_manager.go:58-80) and the UpgradeState* constants
// (consts.go:49-82) while ApplyState's decisions come from the B200 kernel. The same call sequence is exercised,
// compiled, by tests/ through ctypes.
package upgrade

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../.. -lust
#include <stdlib.h>
#include "ust.h"
*/
import "C"

import (
	"context"
	"fmt"
	"unsafe"

	appsv1 "k8s.io/api/apps/v1"
	corev1 "k8s.io/api/core/v1"

	"github.com/NVIDIA/k8s-operator-libs/api/upgrade/v1alpha1"
)

// passOrder is the order in which ApplyState walks the buckets (upgrade_state.go:205-274).
var passOrder = []string{
	UpgradeStateUnknown, UpgradeStateDone, UpgradeStateUpgradeRequired, UpgradeStateCordonRequired,
	UpgradeStateWaitForJobsRequired, UpgradeStatePodDeletionRequired, UpgradeStateDrainRequired,
	UpgradeStateNodeMaintenanceRequired, UpgradeStatePodRestartRequired, UpgradeStateFailed,
	UpgradeStateValidationRequired, UpgradeStateUncordonRequired,
}

var stateCode = map[string]C.uint8_t{
	UpgradeStateUnknown: C.UST_STATE_UNKNOWN, UpgradeStateUpgradeRequired: C.UST_STATE_UPGRADE_REQUIRED,
	UpgradeStateCordonRequired: C.UST_STATE_CORDON_REQUIRED, UpgradeStateWaitForJobsRequired: C.UST_STATE_WAIT_FOR_JOBS_REQUIRED,
	UpgradeStatePodDeletionRequired: C.UST_STATE_POD_DELETION_REQUIRED, UpgradeStateDrainRequired: C.UST_STATE_DRAIN_REQUIRED,
	UpgradeStateNodeMaintenanceRequired: C.UST_STATE_NODE_MAINTENANCE_REQUIRED,
	UpgradeStatePostMaintenanceRequired: C.UST_STATE_POST_MAINTENANCE_REQUIRED,
	UpgradeStatePodRestartRequired: C.UST_STATE_POD_RESTART_REQUIRED, UpgradeStateValidationRequired: C.UST_STATE_VALIDATION_REQUIRED,
	UpgradeStateUncordonRequired: C.UST_STATE_UNCORDON_REQUIRED, UpgradeStateDone: C.UST_STATE_DONE, UpgradeStateFailed: C.UST_STATE_FAILED,
}
var stateName = func() map[C.uint8_t]string {
	m := map[C.uint8_t]string{}
	for k, v := range stateCode {
		m[v] = k
	}
	return m
}()

// Accelerator owns one ust_handle (one GPU). Not re-entrant, like the reference's single reconcile loop
// (node_upgrade_state_provider.go:92-99).
type Accelerator struct{ h *C.ust_handle }

func NewAccelerator(device int) (*Accelerator, error) {
	var h *C.ust_handle
	if rc := C.ust_create(&h, C.int(device)); rc != C.UST_OK {
		return nil, fmt.Errorf("ust_create: %s", C.GoString(C.ust_create_error()))
	}
	return &Accelerator{h: h}, nil
}
func (a *Accelerator) Close() { C.ust_destroy(a.h) }

// flatten DriverUpgradePolicySpec + manager options (upgrade_spec.go:27-110, upgrade_state.go:329-350)
func (m *ClusterUpgradeStateManagerImpl) policy(p *v1alpha1.DriverUpgradePolicySpec) C.ust_policy {
	var c C.ust_policy
	b := func(v bool) C.int32_t {
		if v {
			return 1
		}
		return 0
	}
	c.auto_upgrade = b(p.AutoUpgrade)
	c.max_parallel_upgrades = C.int64_t(p.MaxParallelUpgrades)
	if p.MaxUnavailable != nil {
		// same parsing as intstr.GetScaledValueFromIntOrPercent: Int, "NN%", or error
		c.max_unavailable_kind, c.max_unavailable_value = encodeIntOrPercent(p.MaxUnavailable)
	}
	c.pod_deletion_enabled = b(m.IsPodDeletionEnabled())
	c.validation_enabled = b(m.IsValidationEnabled())
	if p.PodDeletion != nil {
		c.pod_deletion_spec_present, c.pod_deletion_force, c.pod_deletion_delete_emptydir = 1, b(p.PodDeletion.Force), b(p.PodDeletion.DeleteEmptyDir)
	}
	if p.DrainSpec != nil {
		c.drain_enabled, c.drain_force, c.drain_delete_emptydir = b(p.DrainSpec.Enable), b(p.DrainSpec.Force), b(p.DrainSpec.DeleteEmptyDir)
	}
	if p.WaitForCompletion != nil {
		c.wait_selector_set, c.wait_timeout_nonzero = b(p.WaitForCompletion.PodSelector != ""), b(p.WaitForCompletion.TimeoutSecond != 0)
	}
	c.use_maintenance_operator = b(m.opts.Requestor.UseMaintenanceOperator)
	return c
}

// encodeNode evaluates each reference predicate once and packs it (bit meanings: include/ust.h).
func (m *ClusterUpgradeStateManagerImpl) encodeNode(ns *NodeUpgradeState, code C.uint8_t, intern func(string) int32,
	dsIndex func(*NodeUpgradeState) int32) (hot C.uint8_t, flags C.uint32_t, podRev, ds C.int32_t) {
	n := ns.Node
	hot = code
	if m.IsNodeUnschedulable(n) {
		hot |= C.UST_HOT_UNSCHEDULABLE
	}
	if !m.isNodeConditionReady(n) {
		hot |= C.UST_HOT_NOT_READY
	}
	if m.SkipNodeUpgrade(n) {
		hot |= C.UST_HOT_SKIP
	}
	if m.IsUpgradeRequested(n) {
		flags |= C.UST_F_UPGRADE_REQUESTED
	}
	if n.Annotations[GetUpgradeDriverWaitForSafeLoadAnnotationKey()] != "" {
		flags |= C.UST_F_SAFE_LOAD
	}
	if _, ok := n.Annotations[GetUpgradeInitialStateAnnotationKey()]; ok {
		flags |= C.UST_F_INITIAL_STATE_ANNO
	}
	if IsNodeInRequestorMode(n) {
		flags |= C.UST_F_REQUESTOR_MODE
	}
	ds = -1
	if ns.IsOrphanedPod() {
		flags |= C.UST_F_POD_ORPHANED
	} else {
		ds = C.int32_t(dsIndex(ns))
		if hash, err := m.PodManager.GetPodControllerRevisionHash(ns.DriverPod); err != nil {
			hot |= C.UST_HOT_REVISION_HASH_ERROR
		} else {
			podRev = C.int32_t(intern(hash))
		}
	}
	if p := ns.DriverPod; p != nil {
		ready := p.Status.Phase == corev1.PodRunning && len(p.Status.ContainerStatuses) != 0
		for i := range p.Status.ContainerStatuses {
			ready = ready && p.Status.ContainerStatuses[i].Ready
		}
		if ready {
			flags |= C.UST_F_POD_READY
		}
		if m.isDriverPodFailing(p) {
			flags |= C.UST_F_POD_FAILING
		}
		if !p.DeletionTimestamp.IsZero() {
			flags |= C.UST_F_POD_TERMINATING
		}
	}
	if ns.NodeMaintenance != nil {
		flags |= C.UST_F_NM_PRESENT // + UST_F_NM_READY from the Ready condition, upgrade_requestor.go:437-439
	}
	return
}

// ApplyState keeps the reference signature (upgrade_state.go:171-172).
func (m *ClusterUpgradeStateManagerImpl) ApplyStateAccelerated(ctx context.Context, acc *Accelerator,
	currentState *ClusterUpgradeState, upgradePolicy *v1alpha1.DriverUpgradePolicySpec) error {
	if currentState == nil {
		return fmt.Errorf("currentState should not be empty") // upgrade_state.go:175-177
	}
	if upgradePolicy == nil || !upgradePolicy.AutoUpgrade {
		return nil // upgrade_state.go:179-182
	}
	// 1. encode the snapshot bucket by bucket in pass order: SoA index order == replay order, and the
	//    upgrade-required bucket keeps its slice order (upgrade_inplace.go:71).
	var entries []*NodeUpgradeState
	var hot []C.uint8_t
	var flags []C.uint32_t
	var rev, ds []C.int32_t
	// ... (intern table, DaemonSet table with GetDaemonsetControllerRevisionHash per DaemonSet, buckets not in
	//      passOrder appended last so that GetCurrentUnavailableNodes still sees them)
	for _, name := range passOrder {
		for _, ns := range currentState.NodeStates[name] {
			h, f, r, d := m.encodeNode(ns, stateCode[name], nil, nil)
			entries, hot, flags, rev, ds = append(entries, ns), append(hot, h), append(flags, f), append(rev, r), append(ds, d)
		}
	}
	n := len(entries)
	next := make([]C.uint8_t, n)
	actions := make([]C.uint16_t, n)
	var dsRev []C.int32_t
	var counters C.ust_counters
	pol := m.policy(upgradePolicy)
	// 2. one call; Go memory is only borrowed for its duration (cgo pointer rules)
	rc := C.ust_apply_state(acc.h, &pol, C.int64_t(n), (*C.uint8_t)(unsafe.Pointer(&hot[0])), (*C.uint32_t)(unsafe.Pointer(&flags[0])),
		(*C.int32_t)(unsafe.Pointer(&rev[0])), (*C.int32_t)(unsafe.Pointer(&ds[0])), C.int32_t(len(dsRev)), (*C.int32_t)(unsafe.Pointer(&dsRev[0])),
		nil, (*C.uint8_t)(unsafe.Pointer(&next[0])), (*C.uint16_t)(unsafe.Pointer(&actions[0])), nil, &counters)
	// 3. replay through the unchanged L1/L2 interfaces, in order, stopping at the first error exactly like the
	//    sequential loops. On a reference-level abort the kernel already left the unreached nodes untouched.
	var restart []*corev1.Pod
	for i, ns := range entries {
		a := actions[i]
		if a&C.UST_A_ERROR != 0 {
			return fmt.Errorf("%s", C.GoString(C.ust_last_error(acc.h)))
		}
		if a&C.UST_A_CLEAR_UPGRADE_REQUESTED != 0 {
			if err := m.NodeUpgradeStateProvider.ChangeNodeUpgradeAnnotation(ctx, ns.Node, GetUpgradeRequestedAnnotationKey(), "null"); err != nil {
				return err
			}
		}
		if a&C.UST_A_SET_INITIAL_STATE_ANNO != 0 {
			if err := m.NodeUpgradeStateProvider.ChangeNodeUpgradeAnnotation(ctx, ns.Node, GetUpgradeInitialStateAnnotationKey(), trueString); err != nil {
				return err
			}
		}
		if a&C.UST_A_CORDON != 0 {
			if err := m.CordonManager.Cordon(ctx, ns.Node); err != nil {
				return err
			}
		}
		if a&C.UST_A_UNCORDON != 0 {
			if err := m.CordonManager.Uncordon(ctx, ns.Node); err != nil {
				return err
			}
		}
		if a&C.UST_A_UNBLOCK_SAFE_LOAD != 0 {
			if err := m.SafeDriverLoadManager.UnblockLoading(ctx, ns.Node); err != nil {
				return err
			}
		}
		if a&C.UST_A_SET_STATE != 0 {
			if err := m.NodeUpgradeStateProvider.ChangeNodeUpgradeState(ctx, ns.Node, stateName[next[i]]); err != nil {
				return err
			}
		}
		if a&C.UST_A_CLEAR_INITIAL_STATE_ANNO != 0 {
			if err := m.NodeUpgradeStateProvider.ChangeNodeUpgradeAnnotation(ctx, ns.Node, GetUpgradeInitialStateAnnotationKey(), "null"); err != nil {
				return err
			}
		}
		if a&C.UST_A_RESTART_DRIVER_POD != 0 {
			restart = append(restart, ns.DriverPod)
		}
		// UST_A_SCHEDULE_WAIT_CHECK / _POD_EVICTION / _DRAIN: collect node lists and make ONE PodManager /
		// DrainManager call per pass, as common_manager.go:413-414, :443-452, :350-356 do.
	}
	if rc != C.UST_OK {
		return fmt.Errorf("%s", C.GoString(C.ust_last_error(acc.h)))
	}
	return m.PodManager.SchedulePodsRestart(ctx, restart) // common_manager.go:523
}

func encodeIntOrPercent(interface{}) (C.int32_t, C.int64_t) { return C.UST_MAXUNAVAIL_NIL, 0 }

// uid128 parses a Kubernetes UID (a UUID string, "xxxxxxxx-xxxx-xxxx-xxxx-xxxxxxxxxxxx") into the two uint64 the
// ABI joins on; anything that is not 32 hex digits is hashed (FNV-1a) into the same space. (0, 0) is reserved for
// "no owner reference".
func uid128(uid string) (hi, lo uint64) {
	n := 0
	for i := 0; i < len(uid); i++ {
		c := uid[i]
		var v uint64
		switch {
		case c >= '0' && c <= '9':
			v = uint64(c - '0')
		case c >= 'a' && c <= 'f':
			v = uint64(c-'a') + 10
		case c >= 'A' && c <= 'F':
			v = uint64(c-'A') + 10
		case c == '-':
			continue
		default:
			n = -1
		}
		if n < 0 || n >= 32 {
			n = -1
			break
		}
		if n < 16 {
			hi = hi<<4 | v
		} else {
			lo = lo<<4 | v
		}
		n++
	}
	if n != 32 {
		hi, lo = 14695981039346656037, 1099511628211
		for i := 0; i < len(uid); i++ {
			hi = (hi ^ uint64(uid[i])) * 1099511628211
			lo = (lo ^ hi) * 14029467366897019727
		}
	}
	if hi|lo == 0 {
		lo = 1
	}
	return hi, lo
}

// BuildStateAccelerated is the device half of BuildState (upgrade_state.go:99-164): it takes the two API lists
// BuildState already fetched (driver DaemonSets, driver pods) and returns, per pod, the index of the owning
// DaemonSet (-1 orphaned, -2 not a driver pod: dropped), after checking every DaemonSet's pod count against
// DesiredNumberScheduled (upgrade_state.go:128-131). buildNodeUpgradeState (the per-node API Get, :354-378) stays
// in Go and runs for the pods with index >= -1 that are not pending-unscheduled (:149-152).
func (m *ClusterUpgradeStateManagerImpl) BuildStateAccelerated(acc *Accelerator, daemonSets []*appsv1.DaemonSet,
	pods []corev1.Pod, stateCode func(*corev1.Pod) C.uint8_t) ([]int32, C.ust_counters, error) {
	n := len(pods)
	state := make([]C.uint8_t, n+1)
	owner := make([]C.uint64_t, 2*n+2)
	for i := range pods {
		state[i] = stateCode(&pods[i]) // node's upgrade-state label; UST_STATE_EXCLUDED for NodeName=="" && Pending
		if len(pods[i].OwnerReferences) > 0 { // IsOrphanedPod, common_manager.go:225-227
			hi, lo := uid128(string(pods[i].OwnerReferences[0].UID))
			owner[2*i], owner[2*i+1] = C.uint64_t(hi), C.uint64_t(lo)
		}
	}
	dsUID := make([]C.uint64_t, 2*len(daemonSets)+2)
	desired := make([]C.int32_t, len(daemonSets)+1)
	for d, ds := range daemonSets {
		hi, lo := uid128(string(ds.UID))
		dsUID[2*d], dsUID[2*d+1] = C.uint64_t(hi), C.uint64_t(lo)
		desired[d] = C.int32_t(ds.Status.DesiredNumberScheduled)
	}
	idx := make([]int32, n+1)
	var cnt C.ust_counters
	rc := C.ust_build_state_uids(acc.h, C.int64_t(n), &state[0], &owner[0], C.int32_t(len(daemonSets)), &dsUID[0],
		&desired[0], (*C.int32_t)(unsafe.Pointer(&idx[0])), &cnt)
	if rc != C.UST_OK {
		return nil, cnt, fmt.Errorf("%s", C.GoString(C.ust_last_error(acc.h))) // "driver DaemonSet should not have Unscheduled pods"
	}
	return idx[:n], cnt, nil
}

// ApplyStateDelta: for a reconcile loop that tracks resourceVersions. `changed` are the positions (in the order of
// the last full ApplyStateAccelerated call) of the nodes whose Node / Pod / DaemonSet objects changed; only those
// are re-encoded and uploaded, the resident snapshot is evaluated again (ust_apply_state_delta). The replay half is
// the one of ApplyStateAccelerated.
func (a *Accelerator) ApplyStateDelta(pol *C.ust_policy, changed []int64, state []C.uint8_t, flags []C.uint32_t,
	rev, ds []C.int32_t, dsRev []C.int32_t, next []C.uint8_t, actions []C.uint16_t) (C.ust_counters, error) {
	var cnt C.ust_counters
	var idxp *C.int64_t
	var sp *C.uint8_t
	var fp *C.uint32_t
	var rp, dp *C.int32_t
	if len(changed) > 0 {
		idxp, sp, fp = (*C.int64_t)(unsafe.Pointer(&changed[0])), &state[0], &flags[0]
		rp, dp = &rev[0], &ds[0]
	}
	rc := C.ust_apply_state_delta(a.h, pol, C.int64_t(len(changed)), idxp, sp, fp, rp, dp, C.int32_t(len(dsRev)), &dsRev[0],
		&next[0], &actions[0], nil, &cnt)
	if rc != C.UST_OK {
		return cnt, fmt.Errorf("%s", C.GoString(C.ust_last_error(a.h)))
	}
	return cnt, nil
}

// SimulateRollout answers the planning question "how many reconciles does this rollout take under this policy":
// `steps` reconciles on the resident snapshot with ideal actuators (ust_simulate_rollout); history[k] holds the
// counters the k-th reconcile would have reported (GetUpgradesDone / InProgress / Available ...).
func (a *Accelerator) SimulateRollout(pol *C.ust_policy, steps int) ([]C.ust_counters, int, error) {
	history := make([]C.ust_counters, steps+1)
	var done C.int32_t
	rc := C.ust_simulate_rollout(a.h, pol, C.int32_t(steps), &history[0], nil, nil, nil, &done)
	if rc != C.UST_OK {
		return history[:done], int(done), fmt.Errorf("%s", C.GoString(C.ust_last_error(a.h)))
	}
	return history[:steps], int(done), nil
}

// ApplyStatePacked is ApplyState for encoders that intern at most 65535 revision hashes and 127 DaemonSets (every
// real cluster): the two interned columns cross PCIe as uint16 / int8 (ust_apply_state_packed), 8 instead of 13
// bytes per node - the host path is PCIe-bound, so this is the entry point ApplyStateAccelerated should prefer.
func (a *Accelerator) ApplyStatePacked(pol *C.ust_policy, state []C.uint8_t, flags []C.uint32_t, rev16 []C.uint16_t,
	ds8 []C.int8_t, dsRev []C.int32_t, next []C.uint8_t, actions []C.uint16_t) (C.ust_counters, error) {
	var cnt C.ust_counters
	rc := C.ust_apply_state_packed(a.h, pol, C.int64_t(len(state)), &state[0], &flags[0], &rev16[0], &ds8[0],
		C.int32_t(len(dsRev)), &dsRev[0], &next[0], &actions[0], nil, &cnt)
	if rc != C.UST_OK && rc != C.UST_ERR_REVISION_HASH && rc != C.UST_ERR_MAX_UNAVAILABLE && rc != C.UST_ERR_POD_DELETION_SPEC {
		return cnt, fmt.Errorf("%s", C.GoString(C.ust_last_error(a.h)))
	}
	return cnt, nil // reference-level errors come back in cnt.error_code with the outputs cut at the abort point
}
