// upgrade.cpp — see upgrade.hpp. Encode -> ust_apply_state (B200) -> Replay.
#include "upgrade.hpp"

#include <algorithm>
#include <thread>
#include <unordered_map>
#include <cstring>

namespace upgrade {

const char* const UpgradeStateUnknown = "";
const char* const UpgradeStateUpgradeRequired = "upgrade-required";
const char* const UpgradeStateCordonRequired = "cordon-required";
const char* const UpgradeStateWaitForJobsRequired = "wait-for-jobs-required";
const char* const UpgradeStatePodDeletionRequired = "pod-deletion-required";
const char* const UpgradeStateDrainRequired = "drain-required";
const char* const UpgradeStateNodeMaintenanceRequired = "node-maintenance-required";
const char* const UpgradeStatePostMaintenanceRequired = "post-maintenance-required";
const char* const UpgradeStatePodRestartRequired = "pod-restart-required";
const char* const UpgradeStateValidationRequired = "validation-required";
const char* const UpgradeStateUncordonRequired = "uncordon-required";
const char* const UpgradeStateDone = "upgrade-done";
const char* const UpgradeStateFailed = "upgrade-failed";
const char* const PodControllerRevisionHashLabelKey = "controller-revision-hash";

static const char* const kNullString = "null";  // consts.go:90
static const char* const kTrueString = "true";  // consts.go:92
static std::string g_driver_name;                 // util.go:91-99

// The seven keys are functions of the driver name only (util.go:101-133). The reference formats them on every use; an
// encoder that walks a million nodes cannot (six formatted strings per node were a third of Encode's time), so they are
// built once per SetDriverName.
struct DriverKeys {
  bool valid = false;
  std::string state, skip, safeLoad, requested, requestorMode, initialState, waitStart;
};
static DriverKeys g_keys;
static std::string key(const char* tail) { return "nvidia.com/" + g_driver_name + tail; }
static const DriverKeys& keys() {
  if (!g_keys.valid) {
    g_keys.state = key("-driver-upgrade-state");
    g_keys.skip = key("-driver-upgrade.skip");
    g_keys.safeLoad = key("-driver-upgrade.driver-wait-for-safe-load");
    g_keys.requested = key("-driver-upgrade-requested");
    g_keys.requestorMode = key("-driver-upgrade-requestor-mode");
    g_keys.initialState = key("-driver-upgrade.node-initial-state.unschedulable");
    g_keys.waitStart = key("-driver-upgrade-wait-for-pod-completion-start-time");
    g_keys.valid = true;
  }
  return g_keys;
}
void SetDriverName(const std::string& driver) { g_driver_name = driver; g_keys.valid = false; }
std::string GetUpgradeStateLabelKey() { return keys().state; }
std::string GetUpgradeSkipNodeLabelKey() { return keys().skip; }
std::string GetUpgradeDriverWaitForSafeLoadAnnotationKey() { return keys().safeLoad; }
std::string GetUpgradeRequestedAnnotationKey() { return keys().requested; }
std::string GetUpgradeRequestorModeAnnotationKey() { return keys().requestorMode; }
std::string GetUpgradeInitialStateAnnotationKey() { return keys().initialState; }
std::string GetWaitForPodCompletionStartTimeAnnotationKey() { return keys().waitStart; }

bool IsOrphanedPod(const Pod& pod) { return pod.OwnerReferences.empty(); }
bool IsNodeInRequestorMode(const Node& node) { return node.Annotations.count(keys().requestorMode) != 0; }
ClusterUpgradeState NewClusterUpgradeState() { return ClusterUpgradeState(); }

static const char* const kStateNames[13] = {
    UpgradeStateUnknown, UpgradeStateUpgradeRequired, UpgradeStateCordonRequired, UpgradeStateWaitForJobsRequired,
    UpgradeStatePodDeletionRequired, UpgradeStateDrainRequired, UpgradeStateNodeMaintenanceRequired,
    UpgradeStatePostMaintenanceRequired, UpgradeStatePodRestartRequired, UpgradeStateValidationRequired,
    UpgradeStateUncordonRequired, UpgradeStateDone, UpgradeStateFailed};
const char* StateNameOfCode(unsigned code) { return code < 13 ? kStateNames[code] : nullptr; }
int StateCodeOfLabel(const std::string& label) {
  for (int i = 0; i < 13; i++)
    if (label == kStateNames[i]) return i;
  return UST_STATE_OTHER;
}

// buckets in the order ApplyState walks them (upgrade_state.go:205-274)
static const int kPassOrder[12] = {UST_STATE_UNKNOWN, UST_STATE_DONE, UST_STATE_UPGRADE_REQUIRED, UST_STATE_CORDON_REQUIRED,
                                   UST_STATE_WAIT_FOR_JOBS_REQUIRED, UST_STATE_POD_DELETION_REQUIRED, UST_STATE_DRAIN_REQUIRED,
                                   UST_STATE_NODE_MAINTENANCE_REQUIRED, UST_STATE_POD_RESTART_REQUIRED, UST_STATE_FAILED,
                                   UST_STATE_VALIDATION_REQUIRED, UST_STATE_UNCORDON_REQUIRED};

static size_t len(const ClusterUpgradeState& s, const char* name) {
  auto it = s.NodeStates.find(name);
  return it == s.NodeStates.end() ? 0 : it->second.size();
}

// ---- construction -------------------------------------------------------------------------------------------
Error ClusterUpgradeStateManagerImpl::New(int device, StateOptions opts, std::unique_ptr<ClusterUpgradeStateManagerImpl>* out) {
  std::unique_ptr<ClusterUpgradeStateManagerImpl> m(new ClusterUpgradeStateManagerImpl());
  m->opts_ = opts;
  if (ust_create(&m->handle_, device) != UST_OK)
    return Errorf(std::string("failed to create upgrade state manager: ") + ust_create_error());
  *out = std::move(m);
  return std::nullopt;
}
std::unique_ptr<ClusterUpgradeStateManagerImpl> ClusterUpgradeStateManagerImpl::NewDetached(StateOptions opts) {
  std::unique_ptr<ClusterUpgradeStateManagerImpl> m(new ClusterUpgradeStateManagerImpl());
  m->opts_ = opts;
  return m;
}
ClusterUpgradeStateManagerImpl::~ClusterUpgradeStateManagerImpl() { if (handle_) ust_destroy(handle_); }

ClusterUpgradeStateManager& ClusterUpgradeStateManagerImpl::WithPodDeletionEnabled(PodDeletionFilter filter) {
  if (!filter) return *this;  // "Cannot enable PodDeletion state as PodDeletionFilter is nil"  upgrade_state.go:330-333
  filter_ = std::move(filter);
  podDeletionStateEnabled_ = true;
  return *this;
}
ClusterUpgradeStateManager& ClusterUpgradeStateManagerImpl::WithValidationEnabled(const std::string& podSelector) {
  if (podSelector.empty()) return *this;  // upgrade_state.go:342-345
  validationSelector_ = podSelector;
  validationStateEnabled_ = true;
  return *this;
}

// ---- predicates ---------------------------------------------------------------------------------------------
bool ClusterUpgradeStateManagerImpl::IsUpgradeRequested(const Node& n) const {
  auto it = n.Annotations.find(keys().requested);
  return it != n.Annotations.end() && it->second == kTrueString;
}
bool ClusterUpgradeStateManagerImpl::IsNodeUnschedulable(const Node& n) const { return n.Unschedulable; }
bool ClusterUpgradeStateManagerImpl::isNodeConditionReady(const Node& n) const {
  for (const auto& c : n.Conditions)
    if (c.Type == "Ready" && c.Status != "True") return false;
  return true;
}
bool ClusterUpgradeStateManagerImpl::SkipNodeUpgrade(const Node& n) const {
  auto it = n.Labels.find(keys().skip);
  return it != n.Labels.end() && it->second == kTrueString;
}
bool ClusterUpgradeStateManagerImpl::isDriverPodFailing(const Pod& p) const {
  for (const auto& st : p.InitContainerStatuses)
    if (!st.Ready && st.RestartCount > 10) return true;
  for (const auto& st : p.ContainerStatuses)
    if (!st.Ready && st.RestartCount > 10) return true;
  return false;
}

// ---- counters (common_manager.go:146-165, :715-788) ---------------------------------------------------------
int ClusterUpgradeStateManagerImpl::GetTotalManagedNodes(const ClusterUpgradeState& s) const {
  return (int)(len(s, UpgradeStateUnknown) + len(s, UpgradeStateDone) + len(s, UpgradeStateUpgradeRequired) +
               len(s, UpgradeStateCordonRequired) + len(s, UpgradeStateWaitForJobsRequired) +
               len(s, UpgradeStatePodDeletionRequired) + len(s, UpgradeStateFailed) + len(s, UpgradeStateDrainRequired) +
               len(s, UpgradeStatePodRestartRequired) + len(s, UpgradeStateUncordonRequired) +
               len(s, UpgradeStateValidationRequired));
}
int ClusterUpgradeStateManagerImpl::GetUpgradesInProgress(const ClusterUpgradeState& s) const {
  return GetTotalManagedNodes(s) - (int)(len(s, UpgradeStateUnknown) + len(s, UpgradeStateDone) + len(s, UpgradeStateUpgradeRequired));
}
int ClusterUpgradeStateManagerImpl::GetUpgradesDone(const ClusterUpgradeState& s) const { return (int)len(s, UpgradeStateDone); }
int ClusterUpgradeStateManagerImpl::GetUpgradesFailed(const ClusterUpgradeState& s) const { return (int)len(s, UpgradeStateFailed); }
int ClusterUpgradeStateManagerImpl::GetUpgradesPending(const ClusterUpgradeState& s) const { return (int)len(s, UpgradeStateUpgradeRequired); }
int ClusterUpgradeStateManagerImpl::GetCurrentUnavailableNodes(const ClusterUpgradeState& s) const {
  int unavailable = 0;
  for (const auto& kv : s.NodeStates)
    for (const NodeUpgradeState* ns : kv.second) {
      if (IsNodeUnschedulable(*ns->Node)) { unavailable++; continue; }
      if (!isNodeConditionReady(*ns->Node)) unavailable++;
    }
  return unavailable;
}
int ClusterUpgradeStateManagerImpl::GetUpgradesAvailable(const ClusterUpgradeState& s, int maxParallelUpgrades, int maxUnavailable) const {
  const int inProgress = GetUpgradesInProgress(s), total = GetTotalManagedNodes(s);
  int available = maxParallelUpgrades == 0 ? (int)len(s, UpgradeStateUpgradeRequired) : maxParallelUpgrades - inProgress;
  const int currentUnavailable = GetCurrentUnavailableNodes(s) + (int)len(s, UpgradeStateCordonRequired);
  if (available > maxUnavailable) available = maxUnavailable;
  if (currentUnavailable >= maxUnavailable) available = 0;
  else if (maxUnavailable < total && currentUnavailable + available > maxUnavailable) available = maxUnavailable - currentUnavailable;
  return available;
}

// A Kubernetes UID (UUID string) as the two uint64 the ABI joins on: 32 hex digits are taken literally, anything else
// is hashed (FNV-1a) into the same space; (0, 0) is reserved for "no owner reference". Same rule as the Go shim.
struct Uid128 { uint64_t hi, lo; };
static Uid128 uid128(const std::string& uid) {
  Uid128 u{0, 0};
  int n = 0;
  for (char ch : uid) {
    if (ch == '-') continue;
    int v = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : (ch >= 'A' && ch <= 'F') ? ch - 'A' + 10 : -1;
    if (v < 0 || n >= 32) { n = -1; break; }
    if (n < 16) u.hi = (u.hi << 4) | (uint64_t)v; else u.lo = (u.lo << 4) | (uint64_t)v;
    n++;
  }
  if (n != 32) {
    u.hi = 14695981039346656037ull; u.lo = 1099511628211ull;
    for (unsigned char ch : uid) { u.hi = (u.hi ^ ch) * 1099511628211ull; u.lo = (u.lo ^ u.hi) * 14029467366897019727ull; }
  }
  if ((u.hi | u.lo) == 0) u.lo = 1;
  return u;
}

// ---- BuildState (upgrade_state.go:99-164) --------------------------------------------------------------------
Error ClusterUpgradeStateManagerImpl::BuildState(const std::string& ns, const StringMap& driverLabels,
                                                 std::unique_ptr<ClusterUpgradeState>* out) {
  if (handle_ == nullptr) return Errorf("no B200 device bound to this manager: BuildState has no CPU path");
  std::vector<DaemonSet*> dsList;
  if (Error e = K8sClient->ListDaemonSets(ns, driverLabels, &dsList)) return Errorf("error getting DaemonSet list: " + *e);
  std::map<std::string, DaemonSet*> daemonSets;  // UID -> DaemonSet  (common_manager.go:180-186)
  for (DaemonSet* ds : dsList) daemonSets[ds->UID] = ds;
  std::vector<Pod*> podList;
  if (Error e = K8sClient->ListPods(ns, driverLabels, &podList)) return e;

  // The owner join runs on the GPU (ust_build_state_uids): every listed pod goes in with the 128-bit form of its
  // OwnerReferences[0].UID, the DaemonSets with theirs; back comes, per pod, the owning DaemonSet's index, -1 for an
  // orphaned pod, -2 for a pod owned by something else (dropped: GetPodsOwnedbyDs skips it, GetOrphanedPods does
  // not take it - common_manager.go:190-222), after the per-DaemonSet count check of upgrade_state.go:128-131.
  std::vector<DaemonSet*> dsByIndex;
  std::vector<uint64_t> dsUid;
  std::vector<int32_t> desired;
  for (const auto& kv : daemonSets) {
    const Uid128 u = uid128(kv.first);
    dsByIndex.push_back(kv.second);
    dsUid.push_back(u.hi); dsUid.push_back(u.lo);
    desired.push_back(kv.second->DesiredNumberScheduled);
  }
  const size_t np = podList.size();
  std::vector<uint8_t> podState(np + 1);
  std::vector<uint64_t> owner(2 * np + 2, 0);
  std::vector<int32_t> owner_idx(np + 1, -2);
  for (size_t i = 0; i < np; i++) {
    Pod* pod = podList[i];
    // upgrade_state.go:149-152: a pod not yet scheduled to a node is skipped - after the count check
    podState[i] = (pod->NodeName.empty() && pod->Phase == "Pending") ? UST_STATE_EXCLUDED : UST_STATE_OTHER;
    if (!IsOrphanedPod(*pod)) {
      const Uid128 u = uid128(pod->OwnerReferences[0].UID);
      owner[2 * i] = u.hi; owner[2 * i + 1] = u.lo;
    }
  }
  dsUid.resize(dsUid.size() + 2); desired.push_back(0);  // never pass empty vectors' data()
  ust_counters c;
  int rc = ust_build_state_uids(handle_, (int64_t)np, podState.data(), owner.data(), (int32_t)dsByIndex.size(), dsUid.data(),
                                desired.data(), owner_idx.data(), &c);
  if (rc == UST_ERR_DS_UNSCHEDULED) return Errorf("driver DaemonSet should not have Unscheduled pods");  // upgrade_state.go:128-131
  if (rc != UST_OK) return Errorf(ust_last_error(handle_));

  // filteredPodList in the reference's order: DaemonSet by DaemonSet (map order), then the orphans (:126-136)
  std::vector<Pod*> filtered;
  std::vector<uint8_t> state;
  std::vector<std::vector<size_t>> byOwner(dsByIndex.size() + 1);  // last bucket: orphans
  for (size_t i = 0; i < np; i++) {
    if (owner_idx[i] >= 0) byOwner[(size_t)owner_idx[i]].push_back(i);
    else if (owner_idx[i] == -1) byOwner.back().push_back(i);
  }
  for (const auto& bucket : byOwner)
    for (size_t i : bucket) { filtered.push_back(podList[i]); state.push_back(podState[i]); }

  auto st = std::make_unique<ClusterUpgradeState>();
  const std::string labelKey = GetUpgradeStateLabelKey();
  for (size_t i = 0; i < filtered.size(); i++) {
    Pod* pod = filtered[i];
    if (state[i] == UST_STATE_EXCLUDED) continue;
    Node* node = nullptr;
    if (Error e = NodeUpgradeStateProvider->GetNode(pod->NodeName, &node)) return Errorf("unable to get node " + pod->NodeName + ": " + *e);
    auto nus = std::make_unique<NodeUpgradeState>();
    nus->Node = node;
    nus->DriverPod = pod;
    nus->ListIndex = (int64_t)i;
    nus->DriverDaemonSet = IsOrphanedPod(*pod) ? nullptr : daemonSets[pod->OwnerReferences[0].UID];
    if (opts_.Requestor.UseMaintenanceOperator)
      if (Error e = K8sClient->GetNodeMaintenance(node->Name, &nus->NodeMaintenance)) return Errorf("failed while trying to fetch nodeMaintennace obj: " + *e);
    auto it = node->Labels.find(labelKey);
    st->NodeStates[it == node->Labels.end() ? "" : it->second].push_back(nus.get());
    st->owned.push_back(std::move(nus));
  }
  *out = std::move(st);
  return std::nullopt;
}

// ---- ApplyState = Encode -> kernel -> Replay -------------------------------------------------------------------
static void flatten_policy(const DriverUpgradePolicySpec& p, bool podDeletionEnabled, bool validationEnabled, bool useMaintenanceOperator,
                           ust_policy* c) {
  std::memset(c, 0, sizeof(*c));
  c->auto_upgrade = p.AutoUpgrade;
  c->max_parallel_upgrades = p.MaxParallelUpgrades;
  if (p.MaxUnavailable) {
    // intstr.GetScaledValueFromIntOrPercent: Int => IntVal; String must be "<int>%"; anything else is an error
    if (p.MaxUnavailable->Type == IntOrString::Int) {
      c->max_unavailable_kind = UST_MAXUNAVAIL_INT;
      c->max_unavailable_value = p.MaxUnavailable->IntVal;
    } else {
      const std::string& s = p.MaxUnavailable->StrVal;
      bool ok = s.size() >= 2 && s.back() == '%';
      size_t i = ok && (s[0] == '-' || s[0] == '+') ? 1 : 0;
      ok = ok && i < s.size() - 1;
      for (size_t k = i; ok && k + 1 < s.size(); k++) ok = s[k] >= '0' && s[k] <= '9';
      if (ok) {
        c->max_unavailable_kind = UST_MAXUNAVAIL_PERCENT;
        c->max_unavailable_value = std::stoll(s.substr(0, s.size() - 1));
      } else {
        c->max_unavailable_kind = UST_MAXUNAVAIL_INVALID;
      }
    }
  }
  c->pod_deletion_enabled = podDeletionEnabled;
  c->validation_enabled = validationEnabled;
  // a nil PodDeletionSpec is the PodManager's error to raise (pod_manager.go:132-134): the actuator gets the nil
  c->pod_deletion_spec_present = 1;
  if (p.PodDeletion) { c->pod_deletion_force = p.PodDeletion->Force; c->pod_deletion_delete_emptydir = p.PodDeletion->DeleteEmptyDir; }
  if (p.DrainSpec) { c->drain_enabled = p.DrainSpec->Enable; c->drain_force = p.DrainSpec->Force; c->drain_delete_emptydir = p.DrainSpec->DeleteEmptyDir; }
  if (p.WaitForCompletion) { c->wait_selector_set = !p.WaitForCompletion->PodSelector.empty(); c->wait_timeout_nonzero = p.WaitForCompletion->TimeoutSecond != 0; }
  c->use_maintenance_operator = useMaintenanceOperator;
}

// One snapshot entry -> its four SoA values. `ds` / `dsErr`: index of its DaemonSet in the table (-1 = orphaned) and
// whether that DaemonSet's revision-hash lookup failed; `deferred` receives an error the reference would raise when
// its pass reaches the node.
Error ClusterUpgradeStateManagerImpl::encodeOne(const NodeUpgradeState* ns, int code, int32_t ds, bool dsErr,
                                                std::map<std::string, int32_t>* intern, const std::vector<int32_t>& ds_rev,
                                                uint8_t* hot_out, uint32_t* flags_out, int32_t* rev_out, std::string* deferred) {
  auto internHash = [&](const std::string& h) {  // find first: emplace would build (and throw away) a map node per node
    auto it = intern->find(h);
    return it != intern->end() ? it->second : intern->emplace(h, (int32_t)intern->size() + 1).first->second;
  };
  const Node& n = *ns->Node;
  uint8_t hot = (uint8_t)code;
  uint32_t f = 0;
  if (IsNodeUnschedulable(n)) hot |= UST_HOT_UNSCHEDULABLE;
  if (!isNodeConditionReady(n)) hot |= UST_HOT_NOT_READY;
  if (SkipNodeUpgrade(n)) hot |= UST_HOT_SKIP;
  if (IsUpgradeRequested(n)) f |= UST_F_UPGRADE_REQUESTED;
  if (n.Annotations.count(keys().initialState)) f |= UST_F_INITIAL_STATE_ANNO;
  if (IsNodeInRequestorMode(n)) f |= UST_F_REQUESTOR_MODE;
  // ValidationManager.Validate is an actuator with side effects: Replay calls it, at the reference's point in
  // the pass order, and drops the transition when it reports "not done" (common_manager.go:587-596)
  f |= UST_F_VALIDATION_DONE;
  int32_t rev = 0;
  bool synced = false;
  if (ns->IsOrphanedPod()) {
    f |= UST_F_POD_ORPHANED;
  } else {
    std::string podHash;
    if (ns->DriverPod == nullptr || PodManager->GetPodControllerRevisionHash(ns->DriverPod, &podHash) || dsErr)
      hot |= UST_HOT_REVISION_HASH_ERROR;  // pod_manager.go:84-89, :108-110
    else {
      rev = internHash(podHash);
      synced = rev == ds_rev[(size_t)ds];
    }
  }
  // IsWaitingForSafeDriverLoad: the reference consults it in the unknown / upgrade-done passes only
  // (common_manager.go:240) and returns its error there; the pod-restart and validation passes call UnblockLoading
  // unconditionally (:477, :581), a no-op unless the node is waiting - there the predicate only selects whether the
  // call is replayed, and an error from it selects "replay".
  if (code == UST_STATE_UNKNOWN || code == UST_STATE_DONE) {
    bool waiting = false;
    if (Error err = SafeDriverLoadManager->IsWaitingForSafeDriverLoad(&n, &waiting)) {
      if (!(hot & UST_HOT_REVISION_HASH_ERROR)) {  // podInSyncWithDS fails first (:234-238)
        *deferred = *err;
        hot |= UST_HOT_REVISION_HASH_ERROR;         // same abort point: before any action on the node
      }
    } else if (waiting) {
      f |= UST_F_SAFE_LOAD;
    }
  } else if (code == UST_STATE_VALIDATION_REQUIRED || (code == UST_STATE_POD_RESTART_REQUIRED && synced)) {
    bool waiting = false;
    if (SafeDriverLoadManager->IsWaitingForSafeDriverLoad(&n, &waiting) || waiting) f |= UST_F_SAFE_LOAD;
  }
  if (const Pod* p = ns->DriverPod) {
    bool ready = p->Phase == "Running" && !p->ContainerStatuses.empty();  // common_manager.go:617-630
    for (const auto& cs : p->ContainerStatuses) ready = ready && cs.Ready;
    if (ready) f |= UST_F_POD_READY;
    if (isDriverPodFailing(*p)) f |= UST_F_POD_FAILING;
    if (p->DeletionTimestampSet) f |= UST_F_POD_TERMINATING;
  }
  if (ns->NodeMaintenance) {
    f |= UST_F_NM_PRESENT;
    if (ns->NodeMaintenance->ReadyConditionWithReasonReady) f |= UST_F_NM_READY;
  }
  *hot_out = hot;
  *flags_out = f;
  *rev_out = rev;
  return std::nullopt;
}

Error ClusterUpgradeStateManagerImpl::Encode(const ClusterUpgradeState& s, const DriverUpgradePolicySpec& policy, EncodedSnapshot* out) {
  EncodedSnapshot& e = *out;
  e = EncodedSnapshot();
  (void)keys();  // built before any worker thread reads them
  flatten_policy(policy, podDeletionStateEnabled_, validationStateEnabled_, opts_.Requestor.UseMaintenanceOperator, &e.policy);
  // 1. the entries, buckets in pass order: SoA index order == replay order, and the upgrade-required bucket keeps its
  //    slice order
  std::vector<int8_t> codes;
  auto take = [&](const std::vector<NodeUpgradeState*>& bucket, int code) {
    e.entries.insert(e.entries.end(), bucket.begin(), bucket.end());
    codes.insert(codes.end(), bucket.size(), (int8_t)code);
  };
  for (int code : kPassOrder) {
    auto it = s.NodeStates.find(kStateNames[code]);
    if (it != s.NodeStates.end()) take(it->second, code);
  }
  // every other bucket still counts towards GetCurrentUnavailableNodes (common_manager.go:149)
  for (const auto& kv : s.NodeStates) {
    const int code = StateCodeOfLabel(kv.first);
    if (code == UST_STATE_OTHER || code == UST_STATE_POST_MAINTENANCE_REQUIRED) take(kv.second, code);
  }
  const size_t n = e.entries.size();
  // 2. the DaemonSet table: one revision lookup per DaemonSet (pod_manager.go:92-118), in first-use order
  std::map<std::string, int32_t> intern;
  std::map<const DaemonSet*, int32_t> dsIndex;
  std::vector<char> dsHashError;
  e.ds_idx.assign(n, -1);
  for (size_t i = 0; i < n; i++) {
    const NodeUpgradeState* ns = e.entries[i];
    if (ns->IsOrphanedPod()) continue;
    auto it = dsIndex.find(ns->DriverDaemonSet);
    if (it == dsIndex.end()) {
      std::string dsHash;
      const bool bad = (bool)PodManager->GetDaemonsetControllerRevisionHash(ns->DriverDaemonSet, &dsHash);
      it = dsIndex.emplace(ns->DriverDaemonSet, (int32_t)e.ds_rev.size()).first;
      e.ds_rev.push_back(bad ? 0 : intern.emplace(dsHash, (int32_t)intern.size() + 1).first->second);
      dsHashError.push_back(bad ? 1 : 0);
    }
    e.ds_idx[i] = it->second;
  }
  // 3. the nodes, independent of one another. A worker interns the revision hashes it meets in a copy of the table above
  //    (so "pod hash == DaemonSet hash" is an id comparison everywhere); hashes no DaemonSet has get provisional ids that
  //    are made global afterwards.
  e.state.assign(n, 0);
  e.flags.assign(n, 0);
  e.pod_rev.assign(n, 0);
  const int32_t base = (int32_t)intern.size();
  const size_t workers = (size_t)std::max(1, std::min(opts_.EncodeThreads, (int)(n / 4096 + 1)));
  struct Part { std::map<std::string, int32_t> intern; std::vector<std::pair<size_t, std::string>> deferred; Error err; size_t errAt = 0; };
  std::vector<Part> parts(workers);
  auto work = [&](size_t w) {
    Part& p = parts[w];
    p.intern = intern;
    const size_t i0 = n * w / workers, i1 = n * (w + 1) / workers;
    for (size_t i = i0; i < i1; i++) {
      const int32_t ds = e.ds_idx[i];
      uint8_t hot; uint32_t f; int32_t rev; std::string deferred;
      if (Error err = encodeOne(e.entries[i], codes[i], ds, ds >= 0 && dsHashError[(size_t)ds], &p.intern, e.ds_rev, &hot, &f, &rev, &deferred)) {
        p.err = err; p.errAt = i;
        return;
      }
      if (!deferred.empty()) p.deferred.emplace_back(i, deferred);
      e.state[i] = hot; e.flags[i] = f; e.pod_rev[i] = rev;
    }
  };
  if (workers == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (size_t w = 1; w < workers; w++) pool.emplace_back(work, w);
    work(0);
    for (auto& t : pool) t.join();
  }
  for (size_t w = 0; w < workers; w++) {   // first error in index order, like the sequential walk
    if (parts[w].err) return parts[w].err;
    for (auto& d : parts[w].deferred) e.deferred[d.first] = d.second;
  }
  if (workers > 1) {  // provisional ids (> base, per worker) -> global ids
    for (size_t w = 0; w < workers; w++) {
      std::vector<int32_t> remap(parts[w].intern.size() + 1, 0);
      bool moved = false;
      for (const auto& kv : parts[w].intern) {
        if (kv.second <= base) continue;
        const int32_t g = intern.emplace(kv.first, (int32_t)intern.size() + 1).first->second;
        remap[(size_t)kv.second] = g;
        moved = moved || g != kv.second;
      }
      if (!moved) continue;
      const size_t i0 = n * w / workers, i1 = n * (w + 1) / workers;
      for (size_t i = i0; i < i1; i++)
        if (e.pod_rev[i] > base) e.pod_rev[i] = remap[(size_t)e.pod_rev[i]];
    }
  }
  return std::nullopt;
}

Error ClusterUpgradeStateManagerImpl::Replay(const EncodedSnapshot& enc, const DriverUpgradePolicySpec& policy,
                                             const uint8_t* next_state, const uint16_t* actions, int abi_rc,
                                             const ust_counters& counters) {
  const size_t n = enc.entries.size();
  const bool drainEnabled = policy.DrainSpec && policy.DrainSpec->Enable;
  const bool waitSelector = policy.WaitForCompletion && !policy.WaitForCompletion->PodSelector.empty();
  const bool requestor = opts_.Requestor.UseMaintenanceOperator;
  auto setState = [&](size_t i) { return NodeUpgradeStateProvider->ChangeNodeUpgradeState(enc.entries[i]->Node, StateNameOfCode(next_state[i])); };
  auto anno = [&](size_t i, const std::string& k, const char* v) { return NodeUpgradeStateProvider->ChangeNodeUpgradeAnnotation(enc.entries[i]->Node, k, v); };
  auto abortError = [&](long long idx = -1) -> Error {
    if (idx >= 0) { auto it = enc.deferred.find((size_t)idx); if (it != enc.deferred.end()) return Errorf(it->second); }
    return Errorf(ust_last_error(handle_));
  };

  size_t i = 0;
  for (int pass = 0; pass < 12; pass++) {
    const int code = kPassOrder[pass];
    // policy-level abort raised at the start of a pass (intstr parse error, upgrade_inplace.go:54-60)
    if (abi_rc != UST_OK && counters.error_index < 0 && counters.error_pass == pass) return abortError();
    if (code == UST_STATE_NODE_MAINTENANCE_REQUIRED && !requestor) {  // upgrade_state.go:299-309
      while (i < n && (int)(enc.state[i] & UST_HOT_STATE_MASK) == code) i++;
      continue;
    }
    const size_t begin = i;
    std::vector<Node*> batchNodes;
    std::vector<Pod*> restartPods;
    for (; i < n && (int)(enc.state[i] & UST_HOT_STATE_MASK) == code; i++) {
      if (code == UST_STATE_UNCORDON_REQUIRED) continue;  // two sub-passes below
      const unsigned a = actions[i];
      Node* node = enc.entries[i]->Node;
      if (a & UST_A_ERROR) return abortError((long long)i);
      if (a & UST_A_CLEAR_UPGRADE_REQUESTED)
        if (Error e = anno(i, GetUpgradeRequestedAnnotationKey(), kNullString)) return e;
      if (a & UST_A_SET_INITIAL_STATE_ANNO)
        if (Error e = anno(i, GetUpgradeInitialStateAnnotationKey(), kTrueString)) return e;
      if (a & UST_A_CORDON)
        if (Error e = CordonManager->Cordon(node)) return e;
      if (a & UST_A_UNBLOCK_SAFE_LOAD)
        if (Error e = SafeDriverLoadManager->UnblockLoading(node)) return e;
      if (code == UST_STATE_VALIDATION_REQUIRED) {
        bool done = false;
        if (Error e = ValidationManager->Validate(node, &done)) return e;
        if (!done) continue;  // "Validations not complete on the node"
      }
      if ((a & UST_A_NM_CREATE_OR_DELETE) && code == UST_STATE_UPGRADE_REQUIRED)  // upgrade_requestor.go:296
        if (K8sClient != nullptr)
          if (Error e = K8sClient->CreateOrUpdateNodeMaintenance(enc.entries[i])) return e;
      if ((a & UST_A_REQUESTOR_ANNO_CHANGE) && code == UST_STATE_UPGRADE_REQUIRED)
        if (Error e = anno(i, GetUpgradeRequestorModeAnnotationKey(), kTrueString)) return Errorf("failed annotate node for 'upgrade-requestor-mode'. " + *e);
      if (a & UST_A_SET_STATE) {
        Error e = setState(i);
        // common_manager.go:399, :432 deliberately ignore this error in the wait-for-jobs / pod-deletion passes
        if (e && code != UST_STATE_WAIT_FOR_JOBS_REQUIRED && code != UST_STATE_POD_DELETION_REQUIRED) return e;
      }
      if (a & UST_A_CLEAR_INITIAL_STATE_ANNO)
        if (Error e = anno(i, GetUpgradeInitialStateAnnotationKey(), kNullString)) return e;
      if (a & (UST_A_SCHEDULE_WAIT_CHECK | UST_A_SCHEDULE_POD_EVICTION | UST_A_SCHEDULE_DRAIN)) batchNodes.push_back(node);
      if (a & UST_A_RESTART_DRIVER_POD) restartPods.push_back(enc.entries[i]->DriverPod);
    }
    const bool cut = abi_rc != UST_OK && counters.error_pass == pass;  // the kernel stopped inside this pass
    switch (code) {
      case UST_STATE_WAIT_FOR_JOBS_REQUIRED:
        if (waitSelector && !batchNodes.empty()) {  // common_manager.go:404-418
          PodManagerConfig cfg;
          cfg.WaitForCompletionSpec = &*policy.WaitForCompletion;
          cfg.Nodes = batchNodes;
          if (Error e = PodManager->ScheduleCheckOnPodCompletion(cfg)) return e;
        }
        break;
      case UST_STATE_POD_DELETION_REQUIRED:
        if (podDeletionStateEnabled_ && !batchNodes.empty()) {  // common_manager.go:437-452
          PodManagerConfig cfg;
          cfg.DeletionSpec = policy.PodDeletion ? &*policy.PodDeletion : nullptr;
          cfg.DrainEnabled = drainEnabled;
          cfg.Nodes = batchNodes;
          if (Error e = PodManager->SchedulePodEviction(cfg)) return e;
        }
        break;
      case UST_STATE_DRAIN_REQUIRED:
        if (drainEnabled) {  // common_manager.go:346-356 (called even with an empty node list)
          DrainConfiguration cfg;
          cfg.Spec = &*policy.DrainSpec;
          cfg.Nodes = batchNodes;
          if (Error e = DrainManager->ScheduleNodesDrain(cfg)) return e;
        }
        break;
      case UST_STATE_POD_RESTART_REQUIRED:
        if (!cut)  // an abort inside the pass returns before SchedulePodsRestart (common_manager.go:462-523)
          if (Error e = PodManager->SchedulePodsRestart(restartPods)) return e;
        break;
      case UST_STATE_UNCORDON_REQUIRED: {
        // in-place flow first, then the requestor flow (upgrade_state.go:311-325)
        for (size_t k = begin; k < i; k++) {
          if (actions[k] & UST_A_UNCORDON) {
            if (Error e = CordonManager->Uncordon(enc.entries[k]->Node)) return e;
            if (Error e = setState(k)) return e;
          }
        }
        for (size_t k = begin; k < i; k++) {
          if ((actions[k] & UST_A_REQUESTOR_ANNO_CHANGE) && !(actions[k] & UST_A_UNCORDON)) {
            if (Error e = setState(k)) return e;
            if (Error e = anno(k, GetUpgradeRequestorModeAnnotationKey(), kNullString))
              return Errorf("failed to remove '" + GetUpgradeRequestorModeAnnotationKey() + "' annotation . " + *e);
            if (K8sClient != nullptr)
              if (Error e = K8sClient->DeleteOrUpdateNodeMaintenance(enc.entries[k])) return e;  // upgrade_requestor.go:482
          }
        }
      } break;
      default: break;
    }
  }
  if (abi_rc != UST_OK) return abortError(counters.error_index);
  return std::nullopt;
}

Error ClusterUpgradeStateManagerImpl::ApplyState(ClusterUpgradeState* currentState, const DriverUpgradePolicySpec* upgradePolicy) {
  if (currentState == nullptr) return Errorf("currentState should not be empty");  // upgrade_state.go:175-177
  if (upgradePolicy == nullptr || !upgradePolicy->AutoUpgrade) return std::nullopt;  // upgrade_state.go:179-182
  if (handle_ == nullptr) return Errorf("no B200 device bound to this manager: ApplyState has no CPU path");
  EncodedSnapshot enc;
  if (Error e = Encode(*currentState, *upgradePolicy, &enc)) return e;
  const size_t n = enc.entries.size();
  std::vector<uint8_t> next(n + 1);
  std::vector<uint16_t> actions(n + 1);
  enc.state.push_back(0); enc.flags.push_back(0); enc.pod_rev.push_back(0); enc.ds_idx.push_back(0);  // never pass NULL for n == 0
  enc.ds_rev.push_back(0);
  const int rc = ust_apply_state(handle_, &enc.policy, (int64_t)n, enc.state.data(), enc.flags.data(), enc.pod_rev.data(),
                                 enc.ds_idx.data(), (int32_t)enc.ds_rev.size() - 1, enc.ds_rev.data(), nullptr, next.data(),
                                 actions.data(), nullptr, &last_);
  if (rc == UST_ERR_CUDA || rc == UST_ERR_INVALID_ARGUMENT || rc == UST_ERR_NIL_STATE) return Errorf(ust_last_error(handle_));
  return Replay(enc, *upgradePolicy, next.data(), actions.data(), rc, last_);
}

// ---- incremental ApplyState: the resourceVersion-keyed encode cache (upgrade.hpp) --------------------------------------
void ClusterUpgradeStateManagerImpl::ResetIncremental() { cache_ = Cache(); }

int ClusterUpgradeStateManagerImpl::EvaluateCached(const ust_policy& policy, bool full, const std::vector<int64_t>& changed,
                                                   Cache* cache, ust_counters* c) {
  Cache& k = *cache;
  const size_t n = k.slots.size();
  if (handle_ == nullptr) return UST_ERR_CUDA;
  // never pass NULL for empty arrays
  std::vector<int32_t> dsrev = k.ds_rev;
  dsrev.push_back(0);
  if (full) {
    k.next.assign(n + 1, 0);
    k.actions.assign(n + 1, 0);
    std::vector<uint8_t> st = k.state; st.push_back(0);
    std::vector<uint32_t> fl = k.flags; fl.push_back(0);
    std::vector<int32_t> rv = k.pod_rev; rv.push_back(0);
    std::vector<int32_t> di = k.ds_idx; di.push_back(0);
    return ust_apply_state(handle_, &policy, (int64_t)n, st.data(), fl.data(), rv.data(), di.data(), (int32_t)k.ds_rev.size(),
                           dsrev.data(), nullptr, k.next.data(), k.actions.data(), nullptr, c);
  }
  const size_t m = changed.size();
  std::vector<uint8_t> st(m + 1);
  std::vector<uint32_t> fl(m + 1);
  std::vector<int32_t> rv(m + 1), di(m + 1);
  std::vector<int64_t> ix(changed);
  ix.push_back(0);
  for (size_t j = 0; j < m; j++) {
    const size_t i = (size_t)changed[j];
    st[j] = k.state[i]; fl[j] = k.flags[i]; rv[j] = k.pod_rev[i]; di[j] = k.ds_idx[i];
  }
  const int64_t cap = (int64_t)(n / 4 + 1024);
  std::vector<int64_t> oi((size_t)cap + 1);
  std::vector<uint8_t> on((size_t)cap + 1);
  std::vector<uint16_t> oa((size_t)cap + 1);
  int64_t n_out = 0;
  int rc = ust_apply_state_delta_sparse(handle_, &policy, (int64_t)m, ix.data(), st.data(), fl.data(), rv.data(), di.data(),
                                        (int32_t)k.ds_rev.size(), dsrev.data(), cap, oi.data(), on.data(), oa.data(), &n_out, c);
  if (rc == UST_ERR_TRUNCATED) {
    stats_.outputs_received += (int64_t)n;
    return ust_fetch_outputs(handle_, k.next.data(), k.actions.data());
  }
  if (rc == UST_ERR_CUDA || rc == UST_ERR_INVALID_ARGUMENT || rc == UST_ERR_COMM) return rc;
  for (int64_t j = 0; j < n_out; j++) { k.next[(size_t)oi[(size_t)j]] = on[(size_t)j]; k.actions[(size_t)oi[(size_t)j]] = oa[(size_t)j]; }
  stats_.outputs_received += n_out;
  return rc;
}

Error ClusterUpgradeStateManagerImpl::ApplyStateIncremental(ClusterUpgradeState* currentState, const DriverUpgradePolicySpec* upgradePolicy) {
  if (currentState == nullptr) return Errorf("currentState should not be empty");  // upgrade_state.go:175-177
  if (upgradePolicy == nullptr || !upgradePolicy->AutoUpgrade) return std::nullopt;  // upgrade_state.go:179-182
  ust_policy pol;
  flatten_policy(*upgradePolicy, podDeletionStateEnabled_, validationStateEnabled_, opts_.Requestor.UseMaintenanceOperator, &pol);
  for (int attempt = 0; attempt < 2; attempt++) {
    Cache& k = cache_;
    stats_.reconciles += attempt == 0 ? 1 : 0;
    bool full = !k.valid;
    for (auto& sl : k.slots) sl.seen = false;
    // the DaemonSet table: identities are cached by UID, revision hashes are looked up once per DaemonSet per reconcile
    std::map<const DaemonSet*, int32_t> dsOf;
    auto dsIndexOf = [&](const DaemonSet* d) -> int32_t {
      auto it = dsOf.find(d);
      if (it != dsOf.end()) return it->second;
      auto ins = k.dsIndexByUID.emplace(d->UID, (int32_t)k.dsIndexByUID.size());
      const int32_t idx = ins.first->second;
      if ((size_t)idx >= k.ds_rev.size()) { k.ds_rev.resize((size_t)idx + 1, 0); k.dsHashError.resize((size_t)idx + 1, false); }
      std::string dsHash;
      const bool bad = (bool)PodManager->GetDaemonsetControllerRevisionHash(d, &dsHash);
      k.ds_rev[(size_t)idx] = bad ? 0 : k.intern.emplace(dsHash, (int32_t)k.intern.size() + 1).first->second;
      k.dsHashError[(size_t)idx] = bad;
      dsOf.emplace(d, idx);
      return idx;
    };
    std::vector<int64_t> changed;
    // Slots follow BuildState's list order (NodeUpgradeState::ListIndex), of which every bucket's slice order is a
    // subsequence: walk the entries in that order, check that the cached slots still increase along it, append the
    // nodes the cache has not seen. Entries without a ListIndex take their position in the bucket walk instead.
    bool orderBroken = false;
    {
      std::vector<std::pair<int64_t, const NodeUpgradeState*>> order;
      int64_t pos = 0;
      bool haveIndex = true;
      auto collect = [&](const std::vector<NodeUpgradeState*>& v) {
        for (const NodeUpgradeState* ns : v) { haveIndex = haveIndex && ns->ListIndex >= 0; order.emplace_back(ns->ListIndex, ns); pos++; }
      };
      for (int code : kPassOrder) {
        auto it = currentState->NodeStates.find(kStateNames[code]);
        if (it != currentState->NodeStates.end()) collect(it->second);
      }
      for (const auto& kv : currentState->NodeStates) {
        const int code = StateCodeOfLabel(kv.first);
        if (code == UST_STATE_OTHER || code == UST_STATE_POST_MAINTENANCE_REQUIRED) collect(kv.second);
      }
      if (haveIndex) std::stable_sort(order.begin(), order.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
      long long lastSlot = -1;
      for (const auto& o : order) {
        auto ins = k.slotOf.emplace(o.second->Node->Name, k.slots.size());
        if (ins.second) {  // a node the cache has not seen: the snapshot grows
          k.slots.emplace_back();
          k.state.push_back(UST_STATE_EXCLUDED); k.flags.push_back(0); k.pod_rev.push_back(0); k.ds_idx.push_back(-1);
          k.deferredMsg.emplace_back();
          full = true;
        }
        const long long slot = (long long)ins.first->second;
        if (slot < lastSlot) orderBroken = true;
        lastSlot = slot;
      }
    }
    if (orderBroken && attempt == 0) { ResetIncremental(); continue; }  // re-encode in this snapshot's order
    // this reconcile's view in pass order (what Replay walks): entry, its slot
    EncodedSnapshot view;
    view.policy = pol;
    std::vector<size_t> slotOfView;
    auto visit = [&](NodeUpgradeState* ns, int code) -> Error {
      const Node& n = *ns->Node;
      const size_t i = k.slotOf.at(n.Name);
      Cache::Slot& sl = k.slots[i];
      if (sl.seen) return Errorf("node " + n.Name + " appears twice in the snapshot");
      sl.seen = true;
      int32_t ds = -1;
      bool dsErr = false;
      if (!ns->IsOrphanedPod()) { ds = dsIndexOf(ns->DriverDaemonSet); dsErr = k.dsHashError[(size_t)ds]; }
      // everything the encoding of the entry depends on
      const bool versioned = !n.ResourceVersion.empty() && (ns->DriverPod == nullptr || !ns->DriverPod->ResourceVersion.empty());
      std::string sig = std::to_string(code) + "|" + n.ResourceVersion + "|" + (ns->DriverPod ? ns->DriverPod->ResourceVersion : "-") + "|" +
                        std::to_string(ds) + (dsErr ? "!" : "") + "|" + std::to_string(ds >= 0 ? k.ds_rev[(size_t)ds] : 0) + "|" +
                        (ns->NodeMaintenance ? (ns->NodeMaintenance->ReadyConditionWithReasonReady ? "R" : "P") : "-");
      if (!versioned || sig != sl.sig || sl.code != code) {
        uint8_t hot; uint32_t f; int32_t rev; std::string deferred;
        if (Error err = encodeOne(ns, code, ds, dsErr, &k.intern, k.ds_rev, &hot, &f, &rev, &deferred)) return err;
        stats_.encoded++;
        if (hot != k.state[i] || f != k.flags[i] || rev != k.pod_rev[i] || ds != k.ds_idx[i]) {
          k.state[i] = hot; k.flags[i] = f; k.pod_rev[i] = rev; k.ds_idx[i] = ds;
          changed.push_back((int64_t)i);
        }
        k.deferredMsg[i] = deferred;
        sl.sig = versioned ? sig : std::string();
        sl.code = code;
      } else {
        stats_.reused++;
      }
      if (!slotOfView.empty() && (int)(view.state.back() & UST_HOT_STATE_MASK) == code && slotOfView.back() > i)
        orderBroken = true;  // within a bucket, slot order must be the slice order: slots are handed out in it
                             // (upgrade_inplace.go:71) and the first error in it ends the pass
      view.entries.push_back(ns);
      view.state.push_back(k.state[i]);
      slotOfView.push_back(i);
      return std::nullopt;
    };
    Error walkErr;
    for (int code : kPassOrder) {
      auto it = currentState->NodeStates.find(kStateNames[code]);
      if (it == currentState->NodeStates.end()) continue;
      for (NodeUpgradeState* ns : it->second)
        if ((walkErr = visit(ns, code))) break;
      if (walkErr) break;
    }
    if (!walkErr)
      for (const auto& kv : currentState->NodeStates) {
        const int code = StateCodeOfLabel(kv.first);
        if (code != UST_STATE_OTHER && code != UST_STATE_POST_MAINTENANCE_REQUIRED) continue;
        for (NodeUpgradeState* ns : kv.second)
          if ((walkErr = visit(ns, code))) break;
        if (walkErr) break;
      }
    if (walkErr) { ResetIncremental(); return walkErr; }
    if (orderBroken) {  // a bucket's slice order contradicts the cached order (entries without a ListIndex): start over
      ResetIncremental();
      if (attempt == 0) continue;
      return Errorf("incremental ApplyState: a bucket's slice order contradicts the list order");
    }
    // nodes that left the snapshot keep their slot, as "not in snapshot"
    for (size_t i = 0; i < k.slots.size(); i++)
      if (!k.slots[i].seen && (k.state[i] & UST_HOT_STATE_MASK) != UST_STATE_EXCLUDED) {
        k.state[i] = UST_STATE_EXCLUDED; k.flags[i] = 0; k.pod_rev[i] = 0; k.ds_idx[i] = -1;
        k.slots[i].sig.clear(); k.slots[i].code = UST_STATE_EXCLUDED;
        changed.push_back((int64_t)i);
      }
    std::sort(changed.begin(), changed.end());
    if (full) stats_.full_uploads++;
    const int rc = EvaluateCached(pol, full, changed, &k, &last_);
    if (rc == UST_ERR_CUDA || rc == UST_ERR_INVALID_ARGUMENT || rc == UST_ERR_NIL_STATE || rc == UST_ERR_COMM) {
      ResetIncremental();
      return Errorf(handle_ ? ust_last_error(handle_) : "no B200 device bound to this manager: ApplyState has no CPU path");
    }
    k.valid = true;
    // Replay walks this reconcile's view: gather its outputs, translate the abort position
    const size_t nv = view.entries.size();
    std::vector<uint8_t> next(nv + 1);
    std::vector<uint16_t> actions(nv + 1);
    ust_counters c = last_;
    for (size_t v = 0; v < nv; v++) {
      const size_t i = slotOfView[v];
      next[v] = k.next[i];
      actions[v] = k.actions[i];
      if (!k.deferredMsg[i].empty()) view.deferred[v] = k.deferredMsg[i];
      if (last_.error_index == (int64_t)i) c.error_index = (int64_t)v;
    }
    return Replay(view, *upgradePolicy, next.data(), actions.data(), rc, c);
  }
  return Errorf("incremental ApplyState could not settle on a node order");
}

}  // namespace upgrade
