// upgrade.hpp — host-side mirror of github.com/NVIDIA/k8s-operator-libs/pkg/upgrade for the ApplyState /
// BuildState path, written in C++ because the build image has no Go toolchain. Names, argument
// meaning and error behaviour follow the reference so that callers (and tests) read like the Go:
//
//   ClusterUpgradeStateManager      pkg/upgrade/upgrade_state.go:35-53
//   CommonUpgradeStateManager       pkg/upgrade/common_manager.go:23-41
//   NodeUpgradeState / ClusterUpgradeState   common_manager.go:58-80
//   UpgradeState* constants, key getters     consts.go:19-93, util.go:91-155
//   DriverUpgradePolicySpec & sub-specs      api/upgrade/v1alpha1/upgrade_spec.go:27-110
//   actuator interfaces: NodeUpgradeStateProvider (node_upgrade_state_provider.go:33-37), CordonManager
//   (cordon_manager.go:33-36), DrainManager (drain_manager.go:48-50), PodManager (pod_manager.go:53-60),
//   ValidationManager (validation_manager.go:48-50), SafeDriverLoadManager (safe_driver_load_manager.go:74-79)
//
// Decisions are NOT made here: ApplyState encodes the snapshot into the struct-of-arrays of include/ust.h, calls
// ust_apply_state (B200 kernel) and replays the returned per-node action bitmasks through the actuator interfaces
// in the reference's pass order. Without libust.so / a B200 every ApplyState returns an error.
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ust.h"

namespace upgrade {

// ---- errors: Go's `error` ---------------------------------------------------------------------------
using Error = std::optional<std::string>;  // nullopt == nil
inline Error Errorf(std::string s) { return Error(std::move(s)); }

// ---- consts.go:49-82 ----------------------------------------------------------------------------------
extern const char* const UpgradeStateUnknown;
extern const char* const UpgradeStateUpgradeRequired;
extern const char* const UpgradeStateCordonRequired;
extern const char* const UpgradeStateWaitForJobsRequired;
extern const char* const UpgradeStatePodDeletionRequired;
extern const char* const UpgradeStateDrainRequired;
extern const char* const UpgradeStateNodeMaintenanceRequired;
extern const char* const UpgradeStatePostMaintenanceRequired;
extern const char* const UpgradeStatePodRestartRequired;
extern const char* const UpgradeStateValidationRequired;
extern const char* const UpgradeStateUncordonRequired;
extern const char* const UpgradeStateDone;
extern const char* const UpgradeStateFailed;
extern const char* const PodControllerRevisionHashLabelKey;  // pod_manager.go:72

// util.go:91-155
void SetDriverName(const std::string& driver);
std::string GetUpgradeStateLabelKey();
std::string GetUpgradeSkipNodeLabelKey();
std::string GetUpgradeDriverWaitForSafeLoadAnnotationKey();
std::string GetUpgradeRequestedAnnotationKey();
std::string GetUpgradeRequestorModeAnnotationKey();
std::string GetUpgradeInitialStateAnnotationKey();
std::string GetWaitForPodCompletionStartTimeAnnotationKey();

// ---- the slice of corev1 / appsv1 the path reads --------------------------------------------------------
using StringMap = std::map<std::string, std::string>;
struct NodeCondition { std::string Type, Status; };
struct Node {
  std::string Name;
  std::string ResourceVersion;           // metadata.resourceVersion ("" = unknown: the object is re-encoded every reconcile)
  StringMap Labels, Annotations;
  bool Unschedulable = false;            // Spec.Unschedulable
  std::vector<NodeCondition> Conditions;  // Status.Conditions
};
struct ContainerStatus { bool Ready = false; int RestartCount = 0; };
struct OwnerReference { std::string Kind, Name, UID; };
struct Pod {
  std::string Name, Namespace, NodeName;  // Spec.NodeName
  std::string ResourceVersion;
  StringMap Labels;
  std::vector<OwnerReference> OwnerReferences;
  std::string Phase;                      // Status.Phase
  std::vector<ContainerStatus> ContainerStatuses, InitContainerStatuses;
  bool DeletionTimestampSet = false;      // !DeletionTimestamp.IsZero()
};
struct DaemonSet {
  std::string Name, Namespace, UID;
  std::string ResourceVersion;
  int DesiredNumberScheduled = 0;         // Status.DesiredNumberScheduled
};
struct NodeMaintenance {                   // maintenance-operator api v0.3.0, the fields the path reads
  std::string Name;
  bool ReadyConditionWithReasonReady = false;  // upgrade_requestor.go:437-439
};
bool IsOrphanedPod(const Pod& pod);        // common_manager.go:223-225
bool IsNodeInRequestorMode(const Node& node);  // util.go:135-138

// ---- api/upgrade/v1alpha1 --------------------------------------------------------------------------------
struct IntOrString {
  enum Kind { Int, String } Type = Int;
  int64_t IntVal = 0;
  std::string StrVal;
  static IntOrString FromInt(int64_t v) { IntOrString x; x.Type = Int; x.IntVal = v; return x; }
  static IntOrString FromString(std::string s) { IntOrString x; x.Type = String; x.StrVal = std::move(s); return x; }
};
struct WaitForCompletionSpec { std::string PodSelector; int TimeoutSecond = 0; };
struct PodDeletionSpec { bool Force = false; int TimeoutSecond = 300; bool DeleteEmptyDir = false; };
struct DrainSpec { bool Enable = false, Force = false; std::string PodSelector; int TimeoutSecond = 300; bool DeleteEmptyDir = false; };
struct DriverUpgradePolicySpec {
  bool AutoUpgrade = false;
  int64_t MaxParallelUpgrades = 0;
  std::optional<IntOrString> MaxUnavailable;
  std::optional<PodDeletionSpec> PodDeletion;
  std::optional<WaitForCompletionSpec> WaitForCompletion;
  std::optional<upgrade::DrainSpec> DrainSpec;
};

// ---- common_manager.go:58-80 ------------------------------------------------------------------------------
struct NodeUpgradeState {
  upgrade::Node* Node = nullptr;
  Pod* DriverPod = nullptr;
  DaemonSet* DriverDaemonSet = nullptr;
  upgrade::NodeMaintenance* NodeMaintenance = nullptr;
  // Position of the entry's driver pod in BuildState's filteredPodList (upgrade_state.go:126-136), -1 = unknown. Not in
  // the reference's struct: ApplyStateIncremental uses it as the one node order every bucket's slice order is a
  // subsequence of, so that cached entries keep their place when a node changes bucket.
  int64_t ListIndex = -1;
  bool IsOrphanedPod() const { return DriverDaemonSet == nullptr; }
};
struct ClusterUpgradeState {
  std::map<std::string, std::vector<NodeUpgradeState*>> NodeStates;
  std::vector<std::unique_ptr<NodeUpgradeState>> owned;  // BuildState keeps its entries alive here
};
ClusterUpgradeState NewClusterUpgradeState();

// ---- actuator interfaces (stay host-side; driven by the kernel's action bits) --------------------------
struct NodeUpgradeStateProvider {
  virtual ~NodeUpgradeStateProvider() = default;
  virtual Error GetNode(const std::string& nodeName, Node** out) = 0;
  virtual Error ChangeNodeUpgradeState(Node* node, const std::string& newNodeState) = 0;
  virtual Error ChangeNodeUpgradeAnnotation(Node* node, const std::string& key, const std::string& value) = 0;  // "null" deletes
};
struct CordonManager {
  virtual ~CordonManager() = default;
  virtual Error Cordon(Node* node) = 0;
  virtual Error Uncordon(Node* node) = 0;
};
struct DrainConfiguration { const upgrade::DrainSpec* Spec = nullptr; std::vector<Node*> Nodes; };
struct DrainManager {
  virtual ~DrainManager() = default;
  virtual Error ScheduleNodesDrain(const DrainConfiguration& drainConfig) = 0;
};
using PodDeletionFilter = std::function<bool(const Pod&)>;
struct PodManagerConfig {
  std::vector<Node*> Nodes;
  const PodDeletionSpec* DeletionSpec = nullptr;
  const upgrade::WaitForCompletionSpec* WaitForCompletionSpec = nullptr;
  bool DrainEnabled = false;
};
struct PodManager {
  virtual ~PodManager() = default;
  virtual Error ScheduleCheckOnPodCompletion(const PodManagerConfig& config) = 0;
  virtual Error SchedulePodsRestart(const std::vector<Pod*>& pods) = 0;
  virtual Error SchedulePodEviction(const PodManagerConfig& config) = 0;
  virtual PodDeletionFilter GetPodDeletionFilter() = 0;
  virtual Error GetPodControllerRevisionHash(const Pod* pod, std::string* hash) = 0;
  virtual Error GetDaemonsetControllerRevisionHash(const DaemonSet* daemonset, std::string* hash) = 0;
};
struct ValidationManager {
  virtual ~ValidationManager() = default;
  virtual Error Validate(Node* node, bool* done) = 0;
};
struct SafeDriverLoadManager {
  virtual ~SafeDriverLoadManager() = default;
  virtual Error IsWaitingForSafeDriverLoad(const Node* node, bool* waiting) = 0;
  virtual Error UnblockLoading(Node* node) = 0;
};
// what BuildState lists (controller-runtime client in the reference, upgrade_state.go:105-119)
struct K8sClient {
  virtual ~K8sClient() = default;
  virtual Error ListDaemonSets(const std::string& ns, const StringMap& labels, std::vector<DaemonSet*>* out) = 0;
  virtual Error ListPods(const std::string& ns, const StringMap& labels, std::vector<Pod*>* out) = 0;
  virtual Error GetNodeMaintenance(const std::string& nodeName, NodeMaintenance** out) { *out = nullptr; return std::nullopt; }
  // requestor mode: createOrUpdateNodeMaintenance (upgrade_requestor.go:320-368) / deleteOrUpdateNodeMaintenance
  // (:370-414), the NodeMaintenance CRUD of the upgrade-required and uncordon-required passes
  virtual Error CreateOrUpdateNodeMaintenance(NodeUpgradeState* nodeState) { (void)nodeState; return std::nullopt; }
  virtual Error DeleteOrUpdateNodeMaintenance(NodeUpgradeState* nodeState) { (void)nodeState; return std::nullopt; }
};

struct RequestorOptions { bool UseMaintenanceOperator = false; };  // upgrade_requestor.go:527-546 (the switch only)
struct StateOptions {                                                 // upgrade_state.go:94-96
  RequestorOptions Requestor;
  // Not in the reference: host threads Encode may use (<= 1: the calling thread only). Encoding is the host-side cost of a
  // call (object walking, string-keyed map lookups: ~1 us per node) and is independent per node; with more than one thread
  // the injected PodManager::GetPodControllerRevisionHash and SafeDriverLoadManager::IsWaitingForSafeDriverLoad are called
  // concurrently (they are read-only in the reference: pod_manager.go:84-89, safe_driver_load_manager.go:51-57).
  int EncodeThreads = 1;
};

// ---- the encoded snapshot (include/ust.h layout) and its replay ------------------------------------------
struct EncodedSnapshot {
  std::vector<NodeUpgradeState*> entries;  // SoA index -> snapshot entry, buckets in ApplyState's pass order
  std::vector<uint8_t> state;
  std::vector<uint32_t> flags;
  std::vector<int32_t> pod_rev, ds_idx, ds_rev;
  std::map<size_t, std::string> deferred;  // an error the reference raises when it reaches the node (IsWaitingForSafeDriverLoad)
  ust_policy policy{};
};

// ---- common_manager.go:23-41 --------------------------------------------------------------------------------
class CommonUpgradeStateManager {
 public:
  virtual ~CommonUpgradeStateManager() = default;
  virtual int GetTotalManagedNodes(const ClusterUpgradeState& s) const = 0;
  virtual int GetUpgradesInProgress(const ClusterUpgradeState& s) const = 0;
  virtual int GetUpgradesDone(const ClusterUpgradeState& s) const = 0;
  virtual int GetUpgradesAvailable(const ClusterUpgradeState& s, int maxParallelUpgrades, int maxUnavailable) const = 0;
  virtual int GetUpgradesFailed(const ClusterUpgradeState& s) const = 0;
  virtual int GetUpgradesPending(const ClusterUpgradeState& s) const = 0;
  virtual bool IsPodDeletionEnabled() const = 0;
  virtual bool IsValidationEnabled() const = 0;
};

// ---- upgrade_state.go:35-53 -----------------------------------------------------------------------------------
class ClusterUpgradeStateManager : public CommonUpgradeStateManager {
 public:
  virtual ClusterUpgradeStateManager& WithPodDeletionEnabled(PodDeletionFilter filter) = 0;
  virtual ClusterUpgradeStateManager& WithValidationEnabled(const std::string& podSelector) = 0;
  virtual Error BuildState(const std::string& ns, const StringMap& driverLabels, std::unique_ptr<ClusterUpgradeState>* out) = 0;
  virtual Error ApplyState(ClusterUpgradeState* currentState, const DriverUpgradePolicySpec* upgradePolicy) = 0;
};

class ClusterUpgradeStateManagerImpl : public ClusterUpgradeStateManager {
 public:
  // exported fields operators and tests overwrite (common_manager.go:84-100)
  upgrade::K8sClient* K8sClient = nullptr;
  upgrade::DrainManager* DrainManager = nullptr;
  upgrade::PodManager* PodManager = nullptr;
  upgrade::CordonManager* CordonManager = nullptr;
  upgrade::NodeUpgradeStateProvider* NodeUpgradeStateProvider = nullptr;
  upgrade::ValidationManager* ValidationManager = nullptr;
  upgrade::SafeDriverLoadManager* SafeDriverLoadManager = nullptr;

  // NewClusterUpgradeStateManager (upgrade_state.go:65-92): binds CUDA device `device` through ust_create
  static Error New(int device, StateOptions opts, std::unique_ptr<ClusterUpgradeStateManagerImpl>* out);
  // A manager without a device: Encode / Replay work (recording, auditing), ApplyState and BuildState fail loudly.
  static std::unique_ptr<ClusterUpgradeStateManagerImpl> NewDetached(StateOptions opts);
  ~ClusterUpgradeStateManagerImpl() override;

  ClusterUpgradeStateManager& WithPodDeletionEnabled(PodDeletionFilter filter) override;  // upgrade_state.go:329-337
  ClusterUpgradeStateManager& WithValidationEnabled(const std::string& podSelector) override;  // :341-350
  Error BuildState(const std::string& ns, const StringMap& driverLabels, std::unique_ptr<ClusterUpgradeState>* out) override;
  Error ApplyState(ClusterUpgradeState* currentState, const DriverUpgradePolicySpec* upgradePolicy) override;

  int GetTotalManagedNodes(const ClusterUpgradeState& s) const override;
  int GetUpgradesInProgress(const ClusterUpgradeState& s) const override;
  int GetUpgradesDone(const ClusterUpgradeState& s) const override;
  int GetUpgradesAvailable(const ClusterUpgradeState& s, int maxParallelUpgrades, int maxUnavailable) const override;
  int GetUpgradesFailed(const ClusterUpgradeState& s) const override;
  int GetUpgradesPending(const ClusterUpgradeState& s) const override;
  int GetCurrentUnavailableNodes(const ClusterUpgradeState& s) const;
  bool IsPodDeletionEnabled() const override { return podDeletionStateEnabled_; }
  bool IsValidationEnabled() const override { return validationStateEnabled_; }

  // predicates (same names as the Go methods)
  bool IsUpgradeRequested(const Node& n) const;     // common_manager.go:323-325
  bool IsNodeUnschedulable(const Node& n) const;    // :651-653
  bool isNodeConditionReady(const Node& n) const;   // :656-663
  bool SkipNodeUpgrade(const Node& n) const;        // :666-668
  bool isDriverPodFailing(const Pod& p) const;      // :636-648

  // The two halves of ApplyState around the kernel call. Public so that they can be audited separately:
  // Encode evaluates every reference predicate once and fills the struct-of-arrays; Replay performs the calls
  // named by the action bits, in the reference's pass order, stopping at the first error.
  Error Encode(const ClusterUpgradeState& s, const DriverUpgradePolicySpec& policy, EncodedSnapshot* out);
  Error Replay(const EncodedSnapshot& enc, const DriverUpgradePolicySpec& policy, const uint8_t* next_state,
               const uint16_t* actions, int abi_rc, const ust_counters& counters);

  const ust_counters& LastCounters() const { return last_; }

  // ---- incremental ApplyState: the resourceVersion-keyed encode cache -------------------------------
  // The same contract as ApplyState for a reconcile loop that calls it again and again with fresh BuildState
  // snapshots. The manager keeps the encoded snapshot (host and device) from call to call, in a node order that does
  // not move when a node changes bucket (first-seen order = BuildState's pod-list order); an entry is re-encoded only
  // when the resourceVersion of its node, its driver pod or its DaemonSet changed (or is unknown), only the re-encoded
  // entries are uploaded (ust_apply_state_delta_sparse) and only the outputs that differ from the previous reconcile's
  // come back. The reference re-derives everything from node.Labels / Annotations on every reconcile
  // (upgrade_state.go:140-161, common_manager.go:229-604); the provider calls skipped for an unchanged object are
  // GetPodControllerRevisionHash and IsWaitingForSafeDriverLoad, both pure functions of the object in the reference's
  // own implementations (pod_manager.go:84-89, safe_driver_load_manager.go:51-53).
  // Falls back to a full encode + upload when the snapshot grew, when a bucket's slice order no longer follows the
  // cached order (slots are handed out in slice order, upgrade_inplace.go:71), or on the first call.
  Error ApplyStateIncremental(ClusterUpgradeState* currentState, const DriverUpgradePolicySpec* upgradePolicy);
  struct IncrementalStats { int64_t reconciles = 0, full_uploads = 0, encoded = 0, reused = 0, outputs_received = 0; };
  const IncrementalStats& Stats() const { return stats_; }
  void ResetIncremental();

 protected:
  // The device half of ApplyStateIncremental: evaluate the cached snapshot. full: upload all of it and fetch all
  // outputs; else upload the entries `changed` and patch the outputs that differ into cache_.next / cache_.actions.
  // Returns the ABI's return code. (Virtual so that the host-logic test can put the oracle behind the same cache.)
  struct Cache {
    struct Slot { std::string sig; int code = UST_STATE_EXCLUDED; bool seen = false; };
    std::unordered_map<std::string, size_t> slotOf;  // node name -> SoA index
    std::vector<Slot> slots;
    std::vector<uint8_t> state, next;
    std::vector<uint32_t> flags;
    std::vector<int32_t> pod_rev, ds_idx, ds_rev;
    std::vector<uint16_t> actions;
    std::vector<std::string> deferredMsg;            // per slot: the error the reference raises when it reaches the node
    std::map<std::string, int32_t> intern;           // revision hash -> id
    std::map<std::string, int32_t> dsIndexByUID;
    std::vector<bool> dsHashError;
    bool valid = false;
  };
  virtual int EvaluateCached(const ust_policy& policy, bool full, const std::vector<int64_t>& changed, Cache* cache, ust_counters* c);
  ClusterUpgradeStateManagerImpl() = default;

 private:
  Error encodeOne(const NodeUpgradeState* ns, int code, int32_t ds, bool dsErr, std::map<std::string, int32_t>* intern,
                  const std::vector<int32_t>& ds_rev, uint8_t* hot, uint32_t* flags, int32_t* rev, std::string* deferred);
  Cache cache_;
  IncrementalStats stats_;
  ust_handle* handle_ = nullptr;
  StateOptions opts_;
  bool podDeletionStateEnabled_ = false, validationStateEnabled_ = false;
  PodDeletionFilter filter_;
  std::string validationSelector_;
  ust_counters last_{};
};

const char* StateNameOfCode(unsigned code);  // "" for unknown; nullptr for codes without a label
int StateCodeOfLabel(const std::string& label);

}  // namespace upgrade
