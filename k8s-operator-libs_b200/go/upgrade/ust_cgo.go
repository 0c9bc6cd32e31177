//go:build ust

// ust_cgo.go — ApplyState / BuildState of ClusterUpgradeStateManagerImpl on a B200 through libust.so.
//
// Drop this file (and ust_golden_test.go) into pkg/upgrade of github.com/NVIDIA/k8s-operator-libs and build with
// `-tags ust`. It defines, on the reference's own ClusterUpgradeStateManagerImpl,
//
//	func (m *ClusterUpgradeStateManagerImpl) ApplyState(ctx, currentState, upgradePolicy) error
//	func (m *ClusterUpgradeStateManagerImpl) BuildState(ctx, namespace, driverLabels) (*ClusterUpgradeState, error)
//
// so the interface ClusterUpgradeStateManager (upgrade_state.go:35-53), the constructor
// NewClusterUpgradeStateManager (:65-92), the exported fields operators and tests overwrite
// (common_manager.go:84-100) and every UpgradeState* constant stay exactly what they are: gpu-operator /
// network-operator reconcile loops link unchanged. The one edit to the reference's own sources is the rename of its
// two methods, so that both implementations can live in the package (the golden test runs one against the other):
//
//	upgrade_state.go:99   func (m *ClusterUpgradeStateManagerImpl) BuildState(   ->  buildStateReference(
//	upgrade_state.go:171  func (m *ClusterUpgradeStateManagerImpl) ApplyState(   ->  applyStateReference(
//
// (INTEGRATION.md has the two-line patch.) Nothing is decided in Go here: ApplyState encodes the snapshot into the
// struct-of-arrays of include/ust.h (one reference predicate per bit, evaluated with the reference's own helper
// functions), calls ust_apply_state, and replays the per-node action bitmask through the unchanged actuator
// interfaces - NodeUpgradeStateProvider, CordonManager, DrainManager, PodManager, ValidationManager,
// SafeDriverLoadManager - in the reference's pass order, stopping at the first error like the sequential loops do.
// Without libust.so and an sm_100 device every call returns an error: there is no CPU fallback.
//
// This file cannot be compiled in the build image of this repository (no Go toolchain); the same logic, line for
// line, is k8s-operator-libs_b200/host/upgrade.cpp, which is compiled and runs the reference's specs on the GPU.
package upgrade

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../.. -lust -Wl,-rpath,${SRCDIR}/../..
#include <stdint.h>
#include <stdlib.h>
#include "ust.h"
*/
import "C"

import (
	"context"
	"encoding/hex"
	"fmt"
	"os"
	"sort"
	"strconv"
	"strings"
	"sync"
	"unsafe"

	appsv1 "k8s.io/api/apps/v1"
	corev1 "k8s.io/api/core/v1"
	"k8s.io/apimachinery/pkg/api/meta"
	"k8s.io/apimachinery/pkg/types"
	"k8s.io/apimachinery/pkg/util/intstr"
	"sigs.k8s.io/controller-runtime/pkg/client"

	maintenancev1alpha1 "github.com/Mellanox/maintenance-operator/api/v1alpha1"

	"github.com/NVIDIA/k8s-operator-libs/api/upgrade/v1alpha1"
	"github.com/NVIDIA/k8s-operator-libs/pkg/consts"
)

// ---- state codes (include/ust.h, consts.go:49-82) -----------------------------------------------------------------

const (
	ustStateOther    = 13 // any other label value: bucketed by BuildState, counted as unavailable, never processed
	ustStateExcluded = 14
)

// label value of each state code; index = code
var ustStateNames = [13]string{
	UpgradeStateUnknown, UpgradeStateUpgradeRequired, UpgradeStateCordonRequired, UpgradeStateWaitForJobsRequired,
	UpgradeStatePodDeletionRequired, UpgradeStateDrainRequired, UpgradeStateNodeMaintenanceRequired,
	UpgradeStatePostMaintenanceRequired, UpgradeStatePodRestartRequired, UpgradeStateValidationRequired,
	UpgradeStateUncordonRequired, UpgradeStateDone, UpgradeStateFailed,
}

// state code of each Process* pass of ApplyState, in call order (upgrade_state.go:205-274)
var ustPassOrder = [12]int{0, 11, 1, 2, 3, 4, 5, 6, 8, 12, 9, 10}

func ustStateCodeOfLabel(label string) int {
	for code, name := range ustStateNames {
		if name == label {
			return code
		}
	}
	return ustStateOther
}

// ---- the device handle of a manager -----------------------------------------------------------------------------------

type ustHandle struct {
	mu sync.Mutex // ust_handle is not re-entrant; the reconcile loop is single-threaded anyway
	h  *C.ust_handle
}

// one handle per manager, created on first use (the struct itself cannot grow a field from this file)
var ustHandles sync.Map // *ClusterUpgradeStateManagerImpl -> *ustHandle

// UstDeviceEnv names the CUDA device the managers of this process bind to (default 0).
const UstDeviceEnv = "UST_DEVICE"

func (m *ClusterUpgradeStateManagerImpl) ustHandle() (*ustHandle, error) {
	if v, ok := ustHandles.Load(m); ok {
		return v.(*ustHandle), nil
	}
	device := 0
	if s := os.Getenv(UstDeviceEnv); s != "" {
		d, err := strconv.Atoi(s)
		if err != nil {
			return nil, fmt.Errorf("%s=%q is not a device index", UstDeviceEnv, s)
		}
		device = d
	}
	var h *C.ust_handle
	if rc := C.ust_create(&h, C.int(device)); rc != C.UST_OK {
		return nil, fmt.Errorf("ust_create(device %d) failed (%d): %s", device, int(rc), C.GoString(C.ust_create_error()))
	}
	fresh := &ustHandle{h: h}
	if prev, loaded := ustHandles.LoadOrStore(m, fresh); loaded {
		C.ust_destroy(h)
		return prev.(*ustHandle), nil
	}
	return fresh, nil
}

// CloseAccelerator releases the device handle of the manager (optional; handles live as long as the process).
func (m *ClusterUpgradeStateManagerImpl) CloseAccelerator() {
	if v, ok := ustHandles.LoadAndDelete(m); ok {
		uh := v.(*ustHandle)
		uh.mu.Lock()
		C.ust_destroy(uh.h)
		uh.h = nil
		uh.mu.Unlock()
	}
}

// ---- policy -------------------------------------------------------------------------------------------------------------

func ustBool(b bool) C.int32_t {
	if b {
		return 1
	}
	return 0
}

// ustEncodeIntOrPercent mirrors intstr.GetScaledValueFromIntOrPercent's parsing (k8s.io/apimachinery v0.35.1): Int
// => IntVal; String must end in '%' and carry an integer before it; anything else is the error ApplyState returns
// at the upgrade-required pass (upgrade_inplace.go:54-60). The scaling itself (ceil of value * total / 100) needs the
// cluster total and happens on the device.
func ustEncodeIntOrPercent(v *intstr.IntOrString) (kind C.int32_t, value C.int64_t) {
	if v == nil {
		return C.UST_MAXUNAVAIL_NIL, 0 // maxUnavailable = total (upgrade_inplace.go:52)
	}
	switch v.Type {
	case intstr.Int:
		return C.UST_MAXUNAVAIL_INT, C.int64_t(v.IntVal)
	case intstr.String:
		s := v.StrVal
		if !strings.HasSuffix(s, "%") {
			return C.UST_MAXUNAVAIL_INVALID, 0
		}
		n, err := strconv.Atoi(s[:len(s)-1])
		if err != nil {
			return C.UST_MAXUNAVAIL_INVALID, 0
		}
		return C.UST_MAXUNAVAIL_PERCENT, C.int64_t(n)
	}
	return C.UST_MAXUNAVAIL_INVALID, 0
}

func (m *ClusterUpgradeStateManagerImpl) ustFlattenPolicy(p *v1alpha1.DriverUpgradePolicySpec) C.ust_policy {
	var c C.ust_policy
	c.auto_upgrade = ustBool(p.AutoUpgrade)
	c.max_parallel_upgrades = C.int64_t(p.MaxParallelUpgrades)
	c.max_unavailable_kind, c.max_unavailable_value = ustEncodeIntOrPercent(p.MaxUnavailable)
	c.pod_deletion_enabled = ustBool(m.IsPodDeletionEnabled())
	c.validation_enabled = ustBool(m.IsValidationEnabled())
	// a nil PodDeletionSpec is the PodManager's error to raise (pod_manager.go:132-134): the actuator gets the nil
	c.pod_deletion_spec_present = 1
	if p.PodDeletion != nil {
		c.pod_deletion_force = ustBool(p.PodDeletion.Force)
		c.pod_deletion_delete_emptydir = ustBool(p.PodDeletion.DeleteEmptyDir)
	}
	if p.DrainSpec != nil {
		c.drain_enabled = ustBool(p.DrainSpec.Enable)
		c.drain_force = ustBool(p.DrainSpec.Force)
		c.drain_delete_emptydir = ustBool(p.DrainSpec.DeleteEmptyDir)
	}
	if p.WaitForCompletion != nil {
		c.wait_selector_set = ustBool(p.WaitForCompletion.PodSelector != "")
		c.wait_timeout_nonzero = ustBool(p.WaitForCompletion.TimeoutSecond != 0)
	}
	c.use_maintenance_operator = ustBool(m.opts.Requestor.UseMaintenanceOperator)
	return c
}

// ---- encode: ClusterUpgradeState -> struct of arrays -------------------------------------------------------------------

type ustEncoded struct {
	entries  []*NodeUpgradeState // SoA index -> snapshot entry; buckets in ApplyState's pass order
	state    []uint8
	flags    []uint32
	podRev   []int32
	dsIdx    []int32
	dsRev    []int32
	deferred map[int]error // an error the reference raises when it reaches the node (IsWaitingForSafeDriverLoad)
	policy   C.ust_policy
}

func (m *ClusterUpgradeStateManagerImpl) ustEncode(ctx context.Context, s *ClusterUpgradeState,
	policy *v1alpha1.DriverUpgradePolicySpec) *ustEncoded {
	e := &ustEncoded{deferred: map[int]error{}}
	e.policy = m.ustFlattenPolicy(policy)
	intern := map[string]int32{} // revision hash -> small positive id (0 = none)
	internHash := func(h string) int32 {
		if id, ok := intern[h]; ok {
			return id
		}
		id := int32(len(intern) + 1)
		intern[h] = id
		return id
	}
	dsIndex := map[*appsv1.DaemonSet]int32{}
	var dsHashError []bool

	add := func(ns *NodeUpgradeState, code int) {
		node := ns.Node
		hot := uint8(code)
		var f uint32
		if IsNodeUnschedulable(node) { // common_manager.go:651-653
			hot |= C.UST_HOT_UNSCHEDULABLE
		}
		if !m.isNodeConditionReady(node) { // :656-663
			hot |= C.UST_HOT_NOT_READY
		}
		if m.SkipNodeUpgrade(node) { // :666-668
			hot |= C.UST_HOT_SKIP
		}
		if m.IsUpgradeRequested(node) { // :323-325
			f |= C.UST_F_UPGRADE_REQUESTED
		}
		if _, ok := node.Annotations[GetUpgradeInitialStateAnnotationKey()]; ok { // :545, :680
			f |= C.UST_F_INITIAL_STATE_ANNO
		}
		if IsNodeInRequestorMode(node) { // util.go:135-138
			f |= C.UST_F_REQUESTOR_MODE
		}
		// ValidationManager.Validate is an actuator with side effects: replay calls it, at the reference's point in
		// the pass order, and drops the transition when it reports "not done" (common_manager.go:587-596)
		f |= C.UST_F_VALIDATION_DONE

		rev, ds := int32(0), int32(-1)
		synced := false
		if ns.IsOrphanedPod() {
			f |= C.UST_F_POD_ORPHANED
		} else {
			idx, ok := dsIndex[ns.DriverDaemonSet]
			if !ok {
				dsHash, err := m.PodManager.GetDaemonsetControllerRevisionHash(ctx, ns.DriverDaemonSet) // once per DaemonSet
				idx = int32(len(e.dsRev))
				dsIndex[ns.DriverDaemonSet] = idx
				if err != nil {
					e.dsRev = append(e.dsRev, 0)
				} else {
					e.dsRev = append(e.dsRev, internHash(dsHash))
				}
				dsHashError = append(dsHashError, err != nil)
			}
			ds = idx
			podHash, err := "", error(nil)
			if ns.DriverPod == nil {
				err = fmt.Errorf("no driver pod")
			} else {
				podHash, err = m.PodManager.GetPodControllerRevisionHash(ns.DriverPod)
			}
			if err != nil || dsHashError[ds] {
				hot |= C.UST_HOT_REVISION_HASH_ERROR // pod_manager.go:84-89, :108-110
			} else {
				rev = internHash(podHash)
				synced = rev == e.dsRev[ds]
			}
		}
		// IsWaitingForSafeDriverLoad: the reference consults it in the unknown / upgrade-done passes only
		// (common_manager.go:240) and returns its error there; the pod-restart and validation passes call UnblockLoading
		// unconditionally (:477, :581), which is a no-op unless the node is waiting - so there the predicate only
		// selects whether the call is replayed, and an error from it selects "replay".
		switch code {
		case 0, 11:
			waiting, err := m.SafeDriverLoadManager.IsWaitingForSafeDriverLoad(ctx, node)
			if err != nil {
				if hot&C.UST_HOT_REVISION_HASH_ERROR == 0 { // podInSyncWithDS fails first (:234-238)
					e.deferred[len(e.entries)] = err
					hot |= C.UST_HOT_REVISION_HASH_ERROR // same abort point: before any action on the node
				}
			} else if waiting {
				f |= C.UST_F_SAFE_LOAD
			}
		case 8, 9:
			if code == 9 || synced {
				waiting, err := m.SafeDriverLoadManager.IsWaitingForSafeDriverLoad(ctx, node)
				if waiting || err != nil {
					f |= C.UST_F_SAFE_LOAD
				}
			}
		}
		if p := ns.DriverPod; p != nil {
			ready := p.Status.Phase == corev1.PodRunning && len(p.Status.ContainerStatuses) != 0 // :617-630
			for i := range p.Status.ContainerStatuses {
				ready = ready && p.Status.ContainerStatuses[i].Ready
			}
			if ready {
				f |= C.UST_F_POD_READY
			}
			if m.isDriverPodFailing(p) { // :636-648
				f |= C.UST_F_POD_FAILING
			}
			if !p.DeletionTimestamp.IsZero() { // :472
				f |= C.UST_F_POD_TERMINATING
			}
		}
		if ns.NodeMaintenance != nil { // upgrade_requestor.go:420-439
			f |= C.UST_F_NM_PRESENT
			if nm, ok := ns.NodeMaintenance.(*maintenancev1alpha1.NodeMaintenance); ok {
				cond := meta.FindStatusCondition(nm.Status.Conditions, maintenancev1alpha1.ConditionReasonReady)
				if cond != nil && cond.Reason == maintenancev1alpha1.ConditionReasonReady {
					f |= C.UST_F_NM_READY
				}
			}
		}
		e.entries = append(e.entries, ns)
		e.state = append(e.state, hot)
		e.flags = append(e.flags, f)
		e.podRev = append(e.podRev, rev)
		e.dsIdx = append(e.dsIdx, ds)
	}

	// buckets in pass order: SoA index order == replay order, and the upgrade-required bucket keeps its slice order
	// (upgrade_inplace.go:71)
	for _, code := range ustPassOrder {
		for _, ns := range s.NodeStates[ustStateNames[code]] {
			add(ns, code)
		}
	}
	// every other bucket still counts towards GetCurrentUnavailableNodes (common_manager.go:149); sorted for a
	// deterministic encoding
	var others []string
	for label := range s.NodeStates {
		if c := ustStateCodeOfLabel(label); c == ustStateOther || c == 7 {
			others = append(others, label)
		}
	}
	sort.Strings(others)
	for _, label := range others {
		for _, ns := range s.NodeStates[label] {
			add(ns, ustStateCodeOfLabel(label))
		}
	}
	return e
}

// ---- ApplyState (upgrade_state.go:171-281) -------------------------------------------------------------------------------

// ApplyState receives a complete cluster upgrade state and, based on upgrade policy, processes each node's state:
// same contract, guards, call order and error behaviour as the reference method it replaces.
func (m *ClusterUpgradeStateManagerImpl) ApplyState(ctx context.Context,
	currentState *ClusterUpgradeState, upgradePolicy *v1alpha1.DriverUpgradePolicySpec) (err error) {
	m.Log.V(consts.LogLevelInfo).Info("State Manager, got state update")
	if currentState == nil {
		return fmt.Errorf("currentState should not be empty")
	}
	if upgradePolicy == nil || !upgradePolicy.AutoUpgrade {
		m.Log.V(consts.LogLevelInfo).Info("Driver auto upgrade is disabled, skipping")
		return nil
	}
	uh, err := m.ustHandle()
	if err != nil {
		return err
	}
	if m.opts.Requestor.UseMaintenanceOperator {
		if r, ok := m.requestor.(*RequestorNodeStateManagerImpl); ok {
			SetDefaultNodeMaintenance(r.opts, upgradePolicy) // upgrade_requestor.go:283
		}
	}
	enc := m.ustEncode(ctx, currentState, upgradePolicy)
	n := len(enc.entries)
	next := make([]uint8, n+1)
	actions := make([]uint16, n+1)
	var counters C.ust_counters
	// never hand cgo the address of element 0 of an empty slice
	enc.state = append(enc.state, 0)
	enc.flags = append(enc.flags, 0)
	enc.podRev = append(enc.podRev, 0)
	enc.dsIdx = append(enc.dsIdx, 0)
	nDs := len(enc.dsRev)
	enc.dsRev = append(enc.dsRev, 0)

	uh.mu.Lock()
	rc := C.ust_apply_state(uh.h, &enc.policy, C.int64_t(n),
		(*C.uint8_t)(unsafe.Pointer(&enc.state[0])), (*C.uint32_t)(unsafe.Pointer(&enc.flags[0])),
		(*C.int32_t)(unsafe.Pointer(&enc.podRev[0])), (*C.int32_t)(unsafe.Pointer(&enc.dsIdx[0])),
		C.int32_t(nDs), (*C.int32_t)(unsafe.Pointer(&enc.dsRev[0])), nil,
		(*C.uint8_t)(unsafe.Pointer(&next[0])), (*C.uint16_t)(unsafe.Pointer(&actions[0])), nil, &counters)
	lastError := C.GoString(C.ust_last_error(uh.h))
	uh.mu.Unlock()
	switch rc {
	case C.UST_ERR_CUDA, C.UST_ERR_INVALID_ARGUMENT, C.UST_ERR_NIL_STATE, C.UST_ERR_COMM:
		return fmt.Errorf("ust_apply_state failed (%d): %s", int(rc), lastError)
	}
	err = m.ustReplay(ctx, enc, upgradePolicy, next, actions, int(rc), &counters, lastError)
	if err == nil {
		m.Log.V(consts.LogLevelInfo).Info("State Manager, finished processing")
	}
	return err
}

// ustReplay performs the calls named by the action bits, pass by pass in the reference's order, stopping at the first
// error exactly like the sequential loops (upgrade_state.go:205-274).
func (m *ClusterUpgradeStateManagerImpl) ustReplay(ctx context.Context, enc *ustEncoded,
	policy *v1alpha1.DriverUpgradePolicySpec, next []uint8, actions []uint16, abiRC int, counters *C.ust_counters,
	lastError string) error {
	n := len(enc.entries)
	drainEnabled := policy.DrainSpec != nil && policy.DrainSpec.Enable
	waitSelector := policy.WaitForCompletion != nil && policy.WaitForCompletion.PodSelector != ""
	requestorMode := m.opts.Requestor.UseMaintenanceOperator
	provider := m.NodeUpgradeStateProvider
	setState := func(i int) error {
		return provider.ChangeNodeUpgradeState(ctx, enc.entries[i].Node, ustStateNames[next[i]])
	}
	anno := func(i int, key, value string) error {
		return provider.ChangeNodeUpgradeAnnotation(ctx, enc.entries[i].Node, key, value)
	}
	abortError := func(i int) error {
		if i >= 0 {
			if err, ok := enc.deferred[i]; ok {
				return err
			}
		}
		return fmt.Errorf("%s", lastError)
	}
	errorPass, errorIndex := int(counters.error_pass), int(counters.error_index)

	i := 0
	for pass, code := range ustPassOrder {
		// policy-level abort raised at the start of a pass (intstr parse error, upgrade_inplace.go:54-60)
		if abiRC != C.UST_OK && errorIndex < 0 && errorPass == pass {
			return abortError(-1)
		}
		if code == 6 && !requestorMode { // upgrade_state.go:299-309: the in-place flow never touches the bucket
			for i < n && int(enc.state[i]&C.UST_HOT_STATE_MASK) == code {
				i++
			}
			continue
		}
		begin := i
		var batchNodes []*corev1.Node
		restartPods := make([]*corev1.Pod, 0)
		for ; i < n && int(enc.state[i]&C.UST_HOT_STATE_MASK) == code; i++ {
			if code == 10 { // uncordon-required: two sub-passes below
				continue
			}
			a := actions[i]
			ns := enc.entries[i]
			node := ns.Node
			if a&C.UST_A_ERROR != 0 {
				return abortError(i)
			}
			if a&C.UST_A_CLEAR_UPGRADE_REQUESTED != 0 { // upgrade_inplace.go:72-81, upgrade_requestor.go:285-294
				if err := anno(i, GetUpgradeRequestedAnnotationKey(), "null"); err != nil {
					return err
				}
			}
			if a&C.UST_A_SET_INITIAL_STATE_ANNO != 0 { // common_manager.go:253-264
				if err := anno(i, GetUpgradeInitialStateAnnotationKey(), trueString); err != nil {
					return err
				}
			}
			if a&C.UST_A_CORDON != 0 { // :366
				if err := m.CordonManager.Cordon(ctx, node); err != nil {
					return err
				}
			}
			if a&C.UST_A_UNBLOCK_SAFE_LOAD != 0 { // :477, :581
				if err := m.SafeDriverLoadManager.UnblockLoading(ctx, node); err != nil {
					return err
				}
			}
			if code == 9 { // common_manager.go:587-596
				done, err := m.ValidationManager.Validate(ctx, node)
				if err != nil {
					return err
				}
				if !done {
					continue // "Validations not complete on the node"
				}
			}
			if a&C.UST_A_NM_CREATE_OR_DELETE != 0 && code == 1 { // upgrade_requestor.go:296
				r, ok := m.requestor.(*RequestorNodeStateManagerImpl)
				if !ok {
					return fmt.Errorf("requestor mode is enabled but no requestor state manager exists")
				}
				if err := r.createOrUpdateNodeMaintenance(ctx, ns); err != nil {
					return err
				}
			}
			if a&C.UST_A_REQUESTOR_ANNO_CHANGE != 0 && code == 1 { // upgrade_requestor.go:302-306
				if err := anno(i, GetUpgradeRequestorModeAnnotationKey(), trueString); err != nil {
					return fmt.Errorf("failed annotate node for 'upgrade-requestor-mode'. %v", err)
				}
			}
			if a&C.UST_A_SET_STATE != 0 {
				err := setState(i)
				// common_manager.go:399, :432 deliberately ignore this error in the wait-for-jobs / pod-deletion passes
				if err != nil && code != 3 && code != 4 {
					if code == 1 && requestorMode || code == 6 {
						return fmt.Errorf("failed to update node state. %v", err) // upgrade_requestor.go:311, :431, :446
					}
					return err
				}
			}
			if a&C.UST_A_CLEAR_INITIAL_STATE_ANNO != 0 { // common_manager.go:558-565, :699-706
				if err := anno(i, GetUpgradeInitialStateAnnotationKey(), "null"); err != nil {
					return err
				}
			}
			if a&(C.UST_A_SCHEDULE_WAIT_CHECK|C.UST_A_SCHEDULE_POD_EVICTION|C.UST_A_SCHEDULE_DRAIN) != 0 {
				batchNodes = append(batchNodes, node)
			}
			if a&C.UST_A_RESTART_DRIVER_POD != 0 { // :472-474
				restartPods = append(restartPods, ns.DriverPod)
			}
		}
		cut := abiRC != C.UST_OK && errorPass == pass // the kernel stopped inside this pass
		switch code {
		case 3: // common_manager.go:404-418
			if waitSelector && len(batchNodes) != 0 {
				cfg := PodManagerConfig{WaitForCompletionSpec: policy.WaitForCompletion, Nodes: batchNodes}
				if err := m.PodManager.ScheduleCheckOnPodCompletion(ctx, &cfg); err != nil {
					return err
				}
			}
		case 4: // :437-452
			if m.IsPodDeletionEnabled() && len(batchNodes) != 0 {
				cfg := PodManagerConfig{DeletionSpec: policy.PodDeletion, DrainEnabled: drainEnabled, Nodes: batchNodes}
				if err := m.PodManager.SchedulePodEviction(ctx, &cfg); err != nil {
					return err
				}
			}
		case 5: // :346-356 (called even with an empty node list)
			if drainEnabled {
				cfg := DrainConfiguration{Spec: policy.DrainSpec, Nodes: batchNodes}
				if err := m.DrainManager.ScheduleNodesDrain(ctx, &cfg); err != nil {
					return err
				}
			}
		case 8: // an abort inside the pass returns before SchedulePodsRestart (:462-523)
			if !cut {
				if err := m.PodManager.SchedulePodsRestart(ctx, restartPods); err != nil {
					return err
				}
			}
		case 10:
			// in-place flow first, then the requestor flow (upgrade_state.go:311-325)
			for k := begin; k < i; k++ {
				if actions[k]&C.UST_A_UNCORDON != 0 { // upgrade_inplace.go:133-140
					if err := m.CordonManager.Uncordon(ctx, enc.entries[k].Node); err != nil {
						return err
					}
					if err := setState(k); err != nil {
						return err
					}
				}
			}
			for k := begin; k < i; k++ {
				if actions[k]&C.UST_A_REQUESTOR_ANNO_CHANGE != 0 && actions[k]&C.UST_A_UNCORDON == 0 {
					// upgrade_requestor.go:464-485
					if err := setState(k); err != nil {
						return err
					}
					if err := anno(k, GetUpgradeRequestorModeAnnotationKey(), "null"); err != nil {
						return fmt.Errorf("failed to remove '%s' annotation . %v", GetUpgradeRequestorModeAnnotationKey(), err)
					}
					r, ok := m.requestor.(*RequestorNodeStateManagerImpl)
					if !ok {
						return fmt.Errorf("requestor mode is enabled but no requestor state manager exists")
					}
					if err := r.deleteOrUpdateNodeMaintenance(ctx, enc.entries[k]); err != nil {
						return err
					}
				}
			}
		}
	}
	if abiRC != C.UST_OK {
		return abortError(errorIndex)
	}
	return nil
}

// ---- BuildState (upgrade_state.go:99-164) ---------------------------------------------------------------------------------

// ustUID128 is a Kubernetes UID (a UUID string) as the two uint64 the ABI joins on: 32 hex digits are taken literally,
// anything else is hashed (FNV-1a) into the same space; (0, 0) is reserved for "no owner reference". Same rule as
// host/upgrade.cpp.
func ustUID128(uid types.UID) (hi, lo uint64) {
	s := strings.ReplaceAll(string(uid), "-", "")
	if b, err := hex.DecodeString(s); err == nil && len(b) == 16 {
		for k := 0; k < 8; k++ {
			hi = hi<<8 | uint64(b[k])
			lo = lo<<8 | uint64(b[8+k])
		}
	} else {
		hi, lo = 14695981039346656037, 1099511628211
		for k := 0; k < len(uid); k++ {
			hi = (hi ^ uint64(uid[k])) * 1099511628211
			lo = (lo ^ hi) * 14029467366897019727
		}
	}
	if hi|lo == 0 {
		lo = 1
	}
	return hi, lo
}

// BuildState builds a point-in-time snapshot of the driver upgrade state in the cluster: the API lists and the
// per-node GETs are the reference's; the owner join (pod -> DaemonSet by UID), the per-DaemonSet count check against
// DesiredNumberScheduled and the bucket sizes run on the device (ust_build_state_uids).
func (m *ClusterUpgradeStateManagerImpl) BuildState(ctx context.Context, namespace string,
	driverLabels map[string]string) (*ClusterUpgradeState, error) {
	m.Log.V(consts.LogLevelInfo).Info("Building state")
	uh, err := m.ustHandle()
	if err != nil {
		return nil, err
	}
	upgradeState := NewClusterUpgradeState()
	daemonSets, err := m.GetDriverDaemonSets(ctx, namespace, driverLabels)
	if err != nil {
		m.Log.V(consts.LogLevelError).Error(err, "Failed to get driver DaemonSet list")
		return nil, err
	}
	podList := &corev1.PodList{}
	err = m.K8sClient.List(ctx, podList, client.InNamespace(namespace), client.MatchingLabels(driverLabels))
	if err != nil {
		return nil, err
	}

	// DaemonSets in a fixed order (the reference ranges over the map: any order is "the reference's order", G19)
	uids := make([]string, 0, len(daemonSets))
	for uid := range daemonSets {
		uids = append(uids, string(uid))
	}
	sort.Strings(uids)
	dsUID := make([]uint64, 0, 2*len(uids)+2)
	desired := make([]int32, 0, len(uids)+1)
	for _, uid := range uids {
		hi, lo := ustUID128(types.UID(uid))
		dsUID = append(dsUID, hi, lo)
		desired = append(desired, daemonSets[types.UID(uid)].Status.DesiredNumberScheduled)
	}
	np := len(podList.Items)
	podState := make([]uint8, np+1)
	owner := make([]uint64, 2*np+2)
	ownerIdx := make([]int32, np+1)
	for i := range podList.Items {
		pod := &podList.Items[i]
		// upgrade_state.go:149-152: a pod not yet scheduled to a node is skipped - after the count check
		if pod.Spec.NodeName == "" && pod.Status.Phase == corev1.PodPending {
			podState[i] = ustStateExcluded
		} else {
			podState[i] = ustStateOther
		}
		if !IsOrphanedPod(pod) {
			owner[2*i], owner[2*i+1] = ustUID128(pod.OwnerReferences[0].UID)
		}
	}
	dsUID = append(dsUID, 0, 0) // never hand cgo the address of element 0 of an empty slice
	desired = append(desired, 0)
	var counters C.ust_counters
	uh.mu.Lock()
	rc := C.ust_build_state_uids(uh.h, C.int64_t(np), (*C.uint8_t)(unsafe.Pointer(&podState[0])),
		(*C.uint64_t)(unsafe.Pointer(&owner[0])), C.int32_t(len(uids)), (*C.uint64_t)(unsafe.Pointer(&dsUID[0])),
		(*C.int32_t)(unsafe.Pointer(&desired[0])), (*C.int32_t)(unsafe.Pointer(&ownerIdx[0])), &counters)
	lastError := C.GoString(C.ust_last_error(uh.h))
	uh.mu.Unlock()
	if rc == C.UST_ERR_DS_UNSCHEDULED { // upgrade_state.go:128-131
		m.Log.V(consts.LogLevelInfo).Info("Driver DaemonSet has Unscheduled pods", "name",
			daemonSets[types.UID(uids[int(counters.error_index)])].Name)
		return nil, fmt.Errorf("driver DaemonSet should not have Unscheduled pods")
	}
	if rc != C.UST_OK {
		return nil, fmt.Errorf("ust_build_state_uids failed (%d): %s", int(rc), lastError)
	}

	// filteredPodList in the reference's order: DaemonSet by DaemonSet, then the orphans (:126-136)
	byOwner := make([][]int, len(uids)+1) // last bucket: orphans
	for i := 0; i < np; i++ {
		switch {
		case ownerIdx[i] >= 0:
			byOwner[ownerIdx[i]] = append(byOwner[ownerIdx[i]], i)
		case ownerIdx[i] == -1:
			byOwner[len(uids)] = append(byOwner[len(uids)], i)
		}
	}
	upgradeStateLabel := GetUpgradeStateLabelKey()
	for b, bucket := range byOwner {
		for _, i := range bucket {
			pod := &podList.Items[i]
			if podState[i] == ustStateExcluded {
				m.Log.V(consts.LogLevelInfo).Info("Driver Pod has no NodeName, skipping", "pod", pod.Name)
				continue
			}
			var ownerDaemonSet *appsv1.DaemonSet
			if b < len(uids) {
				ownerDaemonSet = daemonSets[types.UID(uids[b])]
			}
			nodeState, err := m.buildNodeUpgradeState(ctx, pod, ownerDaemonSet)
			if err != nil {
				m.Log.V(consts.LogLevelError).Error(err, "Failed to build node upgrade state for pod", "pod", pod)
				return nil, err
			}
			nodeStateLabel := nodeState.Node.Labels[upgradeStateLabel]
			upgradeState.NodeStates[nodeStateLabel] = append(upgradeState.NodeStates[nodeStateLabel], nodeState)
		}
	}
	return &upgradeState, nil
}
