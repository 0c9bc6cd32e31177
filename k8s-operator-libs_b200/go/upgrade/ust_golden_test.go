//go:build ust

// ust_golden_test.go — the reference's own ApplyState against the libust.so-backed one, vector by vector.
//
// The vectors (tests/golden/reference_vectors.json of the ust repository) are the reference's known-answer tests
// for the path, transcribed with file:line provenance from pkg/upgrade/upgrade_state_test.go and
// pod_manager_test.go. For every vector this test builds the ClusterUpgradeState twice, runs
//
//	applyStateReference   the reference's sequential Go loops (upgrade_state.go:171-281, renamed - see ust_cgo.go)
//	ApplyState            encode -> B200 kernel -> replay (ust_cgo.go)
//
// against two copies of the same recording mocks, and requires: the same error / nil, the same node labels,
// annotations and Spec.Unschedulable afterwards, the same sequence of actuator calls, and whatever the Go spec the
// vector was taken from asserts (expect_counts, per-node expect.state). It needs a B200 and libust.so:
//
//	UST_GOLDEN=/path/to/reference_vectors.json go test -tags ust -run TestUstGolden ./pkg/upgrade/
//
// It cannot be run where this repository is built (no Go toolchain); it is the harness SURVEY.md §8(c) promises to
// whoever has one.
package upgrade

import (
	"context"
	"encoding/json"
	"fmt"
	"os"
	"reflect"
	"sort"
	"strings"
	"testing"
	"time"

	"github.com/go-logr/logr"
	appsv1 "k8s.io/api/apps/v1"
	corev1 "k8s.io/api/core/v1"
	metav1 "k8s.io/apimachinery/pkg/apis/meta/v1"
	"k8s.io/apimachinery/pkg/types"
	"k8s.io/apimachinery/pkg/util/intstr"

	maintenancev1alpha1 "github.com/Mellanox/maintenance-operator/api/v1alpha1"

	"github.com/NVIDIA/k8s-operator-libs/api/upgrade/v1alpha1"
)

// ---- the vector file --------------------------------------------------------------------------------------------------

type ustVecPod struct {
	Hash        *string         `json:"hash"`
	Phase       string          `json:"phase"`
	Containers  [][]interface{} `json:"containers"` // [ready bool, restartCount int]
	Init        [][]interface{} `json:"init"`
	Terminating bool            `json:"terminating"`
}

type ustVecNode struct {
	State          string            `json:"state"`
	Unschedulable  bool              `json:"unschedulable"`
	Ready          *string           `json:"ready"`
	Skip           string            `json:"skip"`
	Anno           map[string]string `json:"anno"`
	ValidationDone *bool             `json:"validation_done"`
	DS             bool              `json:"ds"`
	Pod            *ustVecPod        `json:"pod"`
	NM             *struct {
		Ready bool `json:"ready"`
	} `json:"nm"`
	Workload []interface{} `json:"workload"`
	Expect   *struct {
		State string `json:"state"`
	} `json:"expect"`
}

type ustVector struct {
	Name     string          `json:"name"`
	Ref      string          `json:"ref"`
	NilState bool            `json:"nil_state"`
	Policy   json.RawMessage `json:"policy"`
	Options  struct {
		PodDeletionEnabled      bool `json:"podDeletionEnabled"`
		ValidationEnabled       bool `json:"validationEnabled"`
		UseMaintenanceOperator  bool `json:"useMaintenanceOperator"`
	} `json:"options"`
	Nodes            []ustVecNode   `json:"nodes"`
	ExpectCounts     map[string]int `json:"expect_counts"`
	ActuatorError    interface{}    `json:"actuator_error"`
	EvaluateActuator bool           `json:"evaluate_actuators"`
	MockHashGetter   interface{}    `json:"mock_hash_getter"`
}

type ustVecPolicy struct {
	AutoUpgrade         bool                            `json:"autoUpgrade"`
	MaxParallelUpgrades int                             `json:"maxParallelUpgrades"`
	MaxUnavailable      interface{}                     `json:"maxUnavailable"`
	PodDeletion         *v1alpha1.PodDeletionSpec       `json:"podDeletion"`
	Drain               *v1alpha1.DrainSpec             `json:"drain"`
	WaitForCompletion   *v1alpha1.WaitForCompletionSpec `json:"waitForCompletion"`
}

type ustVectorFile struct {
	DaemonsetHash string      `json:"daemonset_hash"`
	ApplyState    []ustVector `json:"apply_state"`
}

// ---- recording mocks with the semantics of the suite's (upgrade_suit_test.go:114-182) ------------------------------------

type ustRecorder struct {
	calls []string
}

func (r *ustRecorder) log(format string, args ...interface{}) {
	r.calls = append(r.calls, fmt.Sprintf(format, args...))
}

func ustNodeNames(nodes []*corev1.Node) string {
	names := make([]string, 0, len(nodes))
	for _, n := range nodes {
		names = append(names, n.Name)
	}
	return strings.Join(names, ",")
}

type ustMockProvider struct{ r *ustRecorder }

func (p *ustMockProvider) GetNode(_ context.Context, name string) (*corev1.Node, error) {
	return nil, fmt.Errorf("GetNode(%s): the golden test never lists", name)
}
func (p *ustMockProvider) ChangeNodeUpgradeState(_ context.Context, node *corev1.Node, state string) error {
	p.r.log("state %s=%s", node.Name, state)
	node.Labels[GetUpgradeStateLabelKey()] = state
	return nil
}
func (p *ustMockProvider) ChangeNodeUpgradeAnnotation(_ context.Context, node *corev1.Node, key, value string) error {
	p.r.log("anno %s %s=%s", node.Name, key, value)
	if value == "null" {
		delete(node.Annotations, key)
	} else {
		node.Annotations[key] = value
	}
	return nil
}

type ustMockCordon struct{ r *ustRecorder }

func (c *ustMockCordon) Cordon(_ context.Context, node *corev1.Node) error {
	c.r.log("cordon %s", node.Name)
	node.Spec.Unschedulable = true
	return nil
}
func (c *ustMockCordon) Uncordon(_ context.Context, node *corev1.Node) error {
	c.r.log("uncordon %s", node.Name)
	node.Spec.Unschedulable = false
	return nil
}

type ustMockDrain struct{ r *ustRecorder }

func (d *ustMockDrain) ScheduleNodesDrain(_ context.Context, cfg *DrainConfiguration) error {
	d.r.log("drain [%s] spec=%+v", ustNodeNames(cfg.Nodes), *cfg.Spec)
	return nil
}

type ustMockPods struct {
	r      *ustRecorder
	dsHash string
	filter PodDeletionFilter
}

func (p *ustMockPods) ScheduleCheckOnPodCompletion(_ context.Context, cfg *PodManagerConfig) error {
	p.r.log("wait-check [%s]", ustNodeNames(cfg.Nodes))
	return nil
}
func (p *ustMockPods) SchedulePodsRestart(_ context.Context, pods []*corev1.Pod) error {
	names := make([]string, 0, len(pods))
	for _, pod := range pods {
		names = append(names, pod.Name)
	}
	p.r.log("restart [%s]", strings.Join(names, ","))
	return nil
}
func (p *ustMockPods) SchedulePodEviction(_ context.Context, cfg *PodManagerConfig) error {
	p.r.log("evict [%s] drain=%v spec-nil=%v", ustNodeNames(cfg.Nodes), cfg.DrainEnabled, cfg.DeletionSpec == nil)
	return nil
}
func (p *ustMockPods) GetPodDeletionFilter() PodDeletionFilter { return p.filter }
func (p *ustMockPods) GetPodControllerRevisionHash(pod *corev1.Pod) (string, error) {
	if hash, ok := pod.Labels[PodControllerRevisionHashLabelKey]; ok {
		return hash, nil
	}
	return "", fmt.Errorf("controller-revision-hash label not present for pod %s", pod.Name)
}
func (p *ustMockPods) GetDaemonsetControllerRevisionHash(_ context.Context, _ *appsv1.DaemonSet) (string, error) {
	return p.dsHash, nil
}

type ustMockValidation struct {
	r    *ustRecorder
	done map[string]bool
}

func (v *ustMockValidation) Validate(_ context.Context, node *corev1.Node) (bool, error) {
	v.r.log("validate %s", node.Name)
	done, ok := v.done[node.Name]
	return done || !ok, nil
}

type ustMockSafeLoad struct{ r *ustRecorder }

func (s *ustMockSafeLoad) IsWaitingForSafeDriverLoad(_ context.Context, node *corev1.Node) (bool, error) {
	return node.Annotations[GetUpgradeDriverWaitForSafeLoadAnnotationKey()] != "", nil
}
func (s *ustMockSafeLoad) UnblockLoading(_ context.Context, node *corev1.Node) error {
	// the real manager is a no-op unless the node is waiting (safe_driver_load_manager.go:57-71): only effective calls
	// are part of the comparison
	if node.Annotations[GetUpgradeDriverWaitForSafeLoadAnnotationKey()] == "" {
		return nil
	}
	s.r.log("unblock %s", node.Name)
	delete(node.Annotations, GetUpgradeDriverWaitForSafeLoadAnnotationKey())
	return nil
}

// ---- building one world from a vector ---------------------------------------------------------------------------------------

type ustWorld struct {
	manager *ClusterUpgradeStateManagerImpl
	state   *ClusterUpgradeState
	nodes   []*corev1.Node
	rec     *ustRecorder
	policy  *v1alpha1.DriverUpgradePolicySpec
}

func ustAnnoKey(short string) string {
	switch short {
	case "upgrade-requested":
		return GetUpgradeRequestedAnnotationKey()
	case "safe-load":
		return GetUpgradeDriverWaitForSafeLoadAnnotationKey()
	case "initial-state":
		return GetUpgradeInitialStateAnnotationKey()
	case "requestor-mode":
		return GetUpgradeRequestorModeAnnotationKey()
	case "wait-start":
		return GetWaitForPodCompletionStartTimeAnnotationKey()
	}
	return short
}

func ustContainerStatuses(rows [][]interface{}) []corev1.ContainerStatus {
	out := make([]corev1.ContainerStatus, 0, len(rows))
	for _, row := range rows {
		ready, _ := row[0].(bool)
		restarts, _ := row[1].(float64)
		out = append(out, corev1.ContainerStatus{Ready: ready, RestartCount: int32(restarts)})
	}
	return out
}

func ustBuildPolicy(raw json.RawMessage) (*v1alpha1.DriverUpgradePolicySpec, error) {
	if len(raw) == 0 || string(raw) == "null" {
		return nil, nil
	}
	var vp ustVecPolicy
	if err := json.Unmarshal(raw, &vp); err != nil {
		return nil, err
	}
	p := &v1alpha1.DriverUpgradePolicySpec{
		AutoUpgrade: vp.AutoUpgrade, MaxParallelUpgrades: vp.MaxParallelUpgrades,
		PodDeletion: vp.PodDeletion, DrainSpec: vp.Drain, WaitForCompletion: vp.WaitForCompletion,
	}
	switch mu := vp.MaxUnavailable.(type) {
	case float64:
		v := intstr.FromInt32(int32(mu))
		p.MaxUnavailable = &v
	case string:
		v := intstr.FromString(mu)
		p.MaxUnavailable = &v
	}
	return p, nil
}

func ustBuildWorld(t *testing.T, v *ustVector, dsHash string) *ustWorld {
	t.Helper()
	rec := &ustRecorder{}
	policy, err := ustBuildPolicy(v.Policy)
	if err != nil {
		t.Fatalf("%s: policy: %v", v.Name, err)
	}
	common := &CommonUpgradeManagerImpl{
		Log:                      logr.Discard(),
		DrainManager:             &ustMockDrain{r: rec},
		CordonManager:            &ustMockCordon{r: rec},
		NodeUpgradeStateProvider: &ustMockProvider{r: rec},
		SafeDriverLoadManager:    &ustMockSafeLoad{r: rec},
	}
	pods := &ustMockPods{r: rec, dsHash: dsHash}
	validation := &ustMockValidation{r: rec, done: map[string]bool{}}
	common.PodManager = pods
	common.ValidationManager = validation
	opts := StateOptions{Requestor: RequestorOptions{UseMaintenanceOperator: v.Options.UseMaintenanceOperator}}
	inplace, err := NewInplaceNodeStateManagerImpl(common)
	if err != nil {
		t.Fatalf("%s: %v", v.Name, err)
	}
	manager := &ClusterUpgradeStateManagerImpl{CommonUpgradeManagerImpl: common, inplace: inplace, opts: opts}
	if v.Options.UseMaintenanceOperator {
		t.Skipf("%s: requestor mode needs an API server for NodeMaintenance objects (envtest); covered by the C++ mirror's specs", v.Name)
	}
	// the With*Enabled options install real actuators (upgrade_state.go:329-350); the mocks stay, the switches are set
	if v.Options.PodDeletionEnabled {
		pods.filter = func(corev1.Pod) bool { return true }
		common.podDeletionStateEnabled = true
	}
	if v.Options.ValidationEnabled {
		common.validationStateEnabled = true
	}

	ds := &appsv1.DaemonSet{ObjectMeta: metav1.ObjectMeta{Name: "driver", Namespace: "default", UID: types.UID("ds-uid")}}
	state := NewClusterUpgradeState()
	w := &ustWorld{manager: manager, state: &state, rec: rec, policy: policy}
	for i := range v.Nodes {
		nd := &v.Nodes[i]
		node := &corev1.Node{ObjectMeta: metav1.ObjectMeta{
			Name: fmt.Sprintf("node-%03d", i), Labels: map[string]string{}, Annotations: map[string]string{}}}
		node.Labels[GetUpgradeStateLabelKey()] = nd.State
		node.Spec.Unschedulable = nd.Unschedulable
		if nd.Ready != nil {
			node.Status.Conditions = []corev1.NodeCondition{{Type: corev1.NodeReady, Status: corev1.ConditionStatus(*nd.Ready)}}
		}
		if nd.Skip != "" {
			node.Labels[GetUpgradeSkipNodeLabelKey()] = nd.Skip
		}
		for k, val := range nd.Anno {
			if k == "wait-start" && strings.HasPrefix(val, "now-") {
				var ago int64
				fmt.Sscanf(val[4:], "%d", &ago)
				val = fmt.Sprintf("%d", time.Now().Unix()-ago)
			}
			node.Annotations[ustAnnoKey(k)] = val
		}
		if nd.ValidationDone != nil {
			validation.done[node.Name] = *nd.ValidationDone
		}
		ns := &NodeUpgradeState{Node: node}
		if nd.DS {
			ns.DriverDaemonSet = ds
		}
		if nd.Pod != nil {
			pod := &corev1.Pod{ObjectMeta: metav1.ObjectMeta{Name: fmt.Sprintf("pod-%03d", i), Namespace: "default",
				Labels: map[string]string{}}}
			if nd.Pod.Hash != nil {
				pod.Labels[PodControllerRevisionHashLabelKey] = *nd.Pod.Hash
			}
			pod.Spec.NodeName = node.Name
			pod.Status.Phase = corev1.PodPhase(nd.Pod.Phase)
			pod.Status.ContainerStatuses = ustContainerStatuses(nd.Pod.Containers)
			pod.Status.InitContainerStatuses = ustContainerStatuses(nd.Pod.Init)
			if nd.Pod.Terminating {
				now := metav1.Now()
				pod.DeletionTimestamp = &now
			}
			ns.DriverPod = pod
		}
		if nd.NM != nil {
			nm := &maintenancev1alpha1.NodeMaintenance{ObjectMeta: metav1.ObjectMeta{Name: "nm-" + node.Name}}
			if nd.NM.Ready {
				nm.Status.Conditions = []metav1.Condition{{Type: maintenancev1alpha1.ConditionReasonReady,
					Reason: maintenancev1alpha1.ConditionReasonReady, Status: metav1.ConditionTrue}}
			}
			ns.NodeMaintenance = nm
		}
		state.NodeStates[nd.State] = append(state.NodeStates[nd.State], ns)
		w.nodes = append(w.nodes, node)
	}
	return w
}

func ustNodeImage(nodes []*corev1.Node) []string {
	out := make([]string, 0, len(nodes))
	for _, n := range nodes {
		keys := make([]string, 0, len(n.Annotations))
		for k, v := range n.Annotations {
			keys = append(keys, k+"="+v)
		}
		sort.Strings(keys)
		out = append(out, fmt.Sprintf("%s state=%q unschedulable=%v anno=%v", n.Name, n.Labels[GetUpgradeStateLabelKey()],
			n.Spec.Unschedulable, keys))
	}
	return out
}

// TestUstGolden: reference ApplyState == accelerated ApplyState on every golden vector.
func TestUstGolden(t *testing.T) {
	path := os.Getenv("UST_GOLDEN")
	if path == "" {
		t.Skip("UST_GOLDEN is not set (path of tests/golden/reference_vectors.json)")
	}
	raw, err := os.ReadFile(path)
	if err != nil {
		t.Fatal(err)
	}
	var file ustVectorFile
	if err := json.Unmarshal(raw, &file); err != nil {
		t.Fatal(err)
	}
	SetDriverName("gpu")
	ctx := context.Background()
	for i := range file.ApplyState {
		v := &file.ApplyState[i]
		t.Run(v.Name, func(t *testing.T) {
			if v.ActuatorError != nil || v.MockHashGetter != nil {
				t.Skip("the vector injects an actuator error through a Go mock the JSON cannot carry")
			}
			ref := ustBuildWorld(t, v, file.DaemonsetHash)
			acc := ustBuildWorld(t, v, file.DaemonsetHash)
			refState, accState := ref.state, acc.state
			if v.NilState {
				refState, accState = nil, nil
			}
			errRef := ref.manager.applyStateReference(ctx, refState, ref.policy)
			errAcc := acc.manager.ApplyState(ctx, accState, acc.policy)
			acc.manager.CloseAccelerator()
			if (errRef == nil) != (errAcc == nil) {
				t.Fatalf("error mismatch: reference %v, accelerated %v", errRef, errAcc)
			}
			if !reflect.DeepEqual(ustNodeImage(ref.nodes), ustNodeImage(acc.nodes)) {
				t.Fatalf("nodes differ:\nreference   %v\naccelerated %v", ustNodeImage(ref.nodes), ustNodeImage(acc.nodes))
			}
			if !reflect.DeepEqual(ref.rec.calls, acc.rec.calls) {
				t.Fatalf("actuator calls differ:\nreference   %v\naccelerated %v", ref.rec.calls, acc.rec.calls)
			}
			// what the Go spec the vector was taken from asserts
			counts := map[string]int{}
			for _, n := range acc.nodes {
				counts[n.Labels[GetUpgradeStateLabelKey()]]++
			}
			for st, want := range v.ExpectCounts {
				if counts[st] != want {
					t.Errorf("%s (%s): %d nodes in %q, the reference's spec expects %d", v.Name, v.Ref, counts[st], st, want)
				}
			}
			for k := range v.Nodes {
				if ex := v.Nodes[k].Expect; ex != nil && ex.State != "" && !v.EvaluateActuator {
					if got := acc.nodes[k].Labels[GetUpgradeStateLabelKey()]; got != ex.State {
						t.Errorf("%s (%s): node %d is %q, the reference's spec expects %q", v.Name, v.Ref, k, got, ex.State)
					}
				}
			}
		})
	}
}
