// The reference's ApplyState specs (pkg/upgrade/upgrade_state_test.go), restated against the C++ mirror of the
// manager. One It() per Go It(); the citation is the Go test it follows. `apply` is how ApplyState is run:
//   * tests/host/upgrade_state_test.cpp  — ClusterUpgradeStateManagerImpl::ApplyState (B200 kernel behind it)
//   * tests/host/host_logic_test.cpp     — Encode -> oracle -> Replay (no GPU: checks the host halves only)
#pragma once
#include <deque>

#include "mocks.hpp"

namespace spec {
using namespace upgrade;
using namespace mocks;

struct Env {
  std::unique_ptr<ClusterUpgradeStateManagerImpl> m;
  NodeUpgradeStateProviderMock provider;
  CordonManagerMock cordon;
  DrainManagerMock drain;
  PodManagerMock pods;
  ValidationManagerMock validation;
  SafeDriverLoadManagerImpl safeLoad{&provider};
  std::deque<Node> nodes;
  std::deque<Pod> podObjs;
  std::deque<NodeUpgradeState> entries;
  DaemonSet daemonSet;
  NodeMaintenance nm;

  void wire() {  // upgrade_state_test.go:54-71
    m->NodeUpgradeStateProvider = &provider;
    m->DrainManager = &drain;
    m->CordonManager = &cordon;
    m->PodManager = &pods;
    m->ValidationManager = &validation;
    m->SafeDriverLoadManager = &safeLoad;
  }
  Node* nodeWithUpgradeState(const std::string& state) {  // upgrade_state_test.go:1788-1795
    nodes.emplace_back();
    nodes.back().Name = "node-" + std::to_string(nodes.size());
    nodes.back().Labels[GetUpgradeStateLabelKey()] = state;
    return &nodes.back();
  }
  Pod* pod(const std::string& hash, const std::string& phase = "", std::vector<ContainerStatus> ctrs = {},
           std::vector<ContainerStatus> init = {}, bool terminating = false, bool hasHash = true) {
    podObjs.emplace_back();
    Pod& p = podObjs.back();
    if (hasHash) p.Labels[PodControllerRevisionHashLabelKey] = hash;
    p.Phase = phase;
    p.ContainerStatuses = std::move(ctrs);
    p.InitContainerStatuses = std::move(init);
    p.DeletionTimestampSet = terminating;
    return &p;
  }
  NodeUpgradeState* entry(Node* n, Pod* p = nullptr, DaemonSet* ds = nullptr, NodeMaintenance* nmObj = nullptr) {
    entries.emplace_back();
    entries.back().Node = n;
    entries.back().DriverPod = p;
    entries.back().DriverDaemonSet = ds;
    entries.back().NodeMaintenance = nmObj;
    return &entries.back();
  }
};

inline std::string getNodeUpgradeState(const Node* n) {
  auto it = n->Labels.find(GetUpgradeStateLabelKey());
  return it == n->Labels.end() ? "" : it->second;
}
inline bool isUnschedulableAnnotationPresent(const Node* n) { return n->Annotations.count(GetUpgradeInitialStateAnnotationKey()) != 0; }

using ApplyFn = std::function<Error(Env&, ClusterUpgradeState*, const DriverUpgradePolicySpec*)>;
using MakeFn = std::function<std::unique_ptr<ClusterUpgradeStateManagerImpl>(StateOptions)>;

inline int countState(const std::vector<NodeUpgradeState*>& v, const char* s) {
  int c = 0;
  for (auto* e : v) c += getNodeUpgradeState(e->Node) == s;
  return c;
}

inline void run(Runner& R, const MakeFn& make, const ApplyFn& apply) {
  SetDriverName("gpu");  // upgrade_suit_test.go:110
  const std::string H = "test-hash-12345", OUT = "test-hash-outdated";
  auto fresh = [&](Env& e, StateOptions o = {}) { e.m = make(o); e.wire(); };
  DriverUpgradePolicySpec AUTO;
  AUTO.AutoUpgrade = true;

  R.it("should fail on nil currentState [upgrade_state_test.go:190-192]", [&] {
    Env e; fresh(e);
    DriverUpgradePolicySpec p;
    EXPECT(R, apply(e, nullptr, &p).has_value());
  });
  R.it("should not fail on nil upgradePolicy [:193-195]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    EXPECT(R, !apply(e, &s, nullptr).has_value());
  });
  R.it("should move up-to-date nodes to Done and outdated nodes to UpgradeRequired [:196-225]", [&] {
    Env e; fresh(e);
    Node *a = e.nodeWithUpgradeState(""), *b = e.nodeWithUpgradeState(""), *c = e.nodeWithUpgradeState(UpgradeStateDone), *d = e.nodeWithUpgradeState(UpgradeStateDone);
    Pod *up = e.pod(H), *out = e.pod(OUT);
    ClusterUpgradeState s;
    s.NodeStates[""] = {e.entry(a, up, &e.daemonSet), e.entry(b, out, &e.daemonSet)};
    s.NodeStates[UpgradeStateDone] = {e.entry(c, up, &e.daemonSet), e.entry(d, out, &e.daemonSet)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(a) == UpgradeStateDone);
    EXPECT(R, getNodeUpgradeState(b) == UpgradeStateUpgradeRequired);
    EXPECT(R, getNodeUpgradeState(c) == UpgradeStateDone);
    EXPECT(R, getNodeUpgradeState(d) == UpgradeStateUpgradeRequired);
  });
  R.it("should annotate unschedulable outdated nodes with the initial state [:226-270]", [&] {
    Env e; fresh(e);
    Node *a = e.nodeWithUpgradeState(""), *b = e.nodeWithUpgradeState(""), *c = e.nodeWithUpgradeState(UpgradeStateDone), *d = e.nodeWithUpgradeState(UpgradeStateDone);
    b->Unschedulable = d->Unschedulable = true;
    Pod *up = e.pod(H), *out = e.pod(OUT);
    ClusterUpgradeState s;
    s.NodeStates[""] = {e.entry(a, up, &e.daemonSet), e.entry(b, out, &e.daemonSet)};
    s.NodeStates[UpgradeStateDone] = {e.entry(c, up, &e.daemonSet), e.entry(d, out, &e.daemonSet)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(b) == UpgradeStateUpgradeRequired && getNodeUpgradeState(d) == UpgradeStateUpgradeRequired);
    EXPECT(R, isUnschedulableAnnotationPresent(b) && isUnschedulableAnnotationPresent(d));
    EXPECT(R, !isUnschedulableAnnotationPresent(a) && !isUnschedulableAnnotationPresent(c));
  });
  R.it("should move up-to-date nodes with the safe driver loading annotation to UpgradeRequired [:271-293]", [&] {
    Env e; fresh(e);
    Node* n = e.nodeWithUpgradeState(UpgradeStateDone);
    n->Annotations[GetUpgradeDriverWaitForSafeLoadAnnotationKey()] = "true";
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateDone] = {e.entry(n, e.pod(H), &e.daemonSet)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(n) == UpgradeStateUpgradeRequired);
  });
  auto upgradeRequired = [&](Env& e, int n, int unschedulableLast = 0) {
    std::vector<NodeUpgradeState*> v;
    for (int i = 0; i < n; i++) {
      Node* node = e.nodeWithUpgradeState(UpgradeStateUpgradeRequired);
      if (i >= n - unschedulableLast) node->Unschedulable = true;
      v.push_back(e.entry(node));
    }
    return v;
  };
  R.it("should schedule upgrade on all nodes if maxParallel upgrades is set to 0 [:294-320]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateUpgradeRequired] = upgradeRequired(e, 5);
    DriverUpgradePolicySpec p = AUTO;
    EXPECT(R, !apply(e, &s, &p).has_value());
    EXPECT(R, countState(s.NodeStates[UpgradeStateUpgradeRequired], UpgradeStateCordonRequired) == 5);
  });
  R.it("should start upgrade on limited amount of nodes if maxParallel < node count [:321-349]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateUpgradeRequired] = upgradeRequired(e, 5);
    DriverUpgradePolicySpec p = AUTO;
    p.MaxParallelUpgrades = 3;
    EXPECT(R, !apply(e, &s, &p).has_value());
    EXPECT(R, countState(s.NodeStates[UpgradeStateUpgradeRequired], UpgradeStateUpgradeRequired) == 2);
    EXPECT(R, countState(s.NodeStates[UpgradeStateUpgradeRequired], UpgradeStateCordonRequired) == 3);
  });
  R.it("should start additional upgrades if maxParallelUpgrades limit is not reached [:350-383]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateUpgradeRequired] = upgradeRequired(e, 2);
    for (int i = 0; i < 3; i++) s.NodeStates[UpgradeStateCordonRequired].push_back(e.entry(e.nodeWithUpgradeState(UpgradeStateCordonRequired)));
    DriverUpgradePolicySpec p = AUTO;
    p.MaxParallelUpgrades = 4;
    p.DrainSpec = DrainSpec{};
    p.DrainSpec->Enable = true;
    EXPECT(R, !apply(e, &s, &p).has_value());
    auto all = s.NodeStates[UpgradeStateUpgradeRequired];
    all.insert(all.end(), s.NodeStates[UpgradeStateCordonRequired].begin(), s.NodeStates[UpgradeStateCordonRequired].end());
    EXPECT(R, countState(all, UpgradeStateUpgradeRequired) == 1);
    EXPECT(R, countState(all, UpgradeStateCordonRequired) + countState(all, UpgradeStateWaitForJobsRequired) == 4);
  });
  R.it("maxParallel 0 and maxUnavailable 100% schedules all [:384-412]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateUpgradeRequired] = upgradeRequired(e, 5, 2);
    DriverUpgradePolicySpec p = AUTO;
    p.MaxUnavailable = IntOrString::FromString("100%");
    EXPECT(R, !apply(e, &s, &p).has_value());
    EXPECT(R, countState(s.NodeStates[UpgradeStateUpgradeRequired], UpgradeStateCordonRequired) == 5);
  });
  R.it("maxUnavailable 50% with two nodes already cordoned [:413-440]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateUpgradeRequired] = upgradeRequired(e, 5, 2);
    DriverUpgradePolicySpec p = AUTO;
    p.MaxUnavailable = IntOrString::FromString("50%");
    EXPECT(R, !apply(e, &s, &p).has_value());
    EXPECT(R, countState(s.NodeStates[UpgradeStateUpgradeRequired], UpgradeStateUpgradeRequired) == 2);
    EXPECT(R, countState(s.NodeStates[UpgradeStateUpgradeRequired], UpgradeStateCordonRequired) == 3);
  });
  R.it("50% maxUnavailable with some unavailable nodes already upgraded [:441-513]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateUpgradeRequired] = upgradeRequired(e, 3);
    Pod* up = e.pod(H, "Running");
    for (int i = 0; i < 2; i++) {
      Node* n = e.nodeWithUpgradeState(UpgradeStateDone);
      n->Unschedulable = true;
      s.NodeStates[UpgradeStateDone].push_back(e.entry(n, up, &e.daemonSet));
    }
    DriverUpgradePolicySpec p = AUTO;
    p.MaxUnavailable = IntOrString::FromString("50%");
    EXPECT(R, !apply(e, &s, &p).has_value());
    EXPECT(R, e.pods.restarted.empty());
    EXPECT(R, countState(s.NodeStates[UpgradeStateDone], UpgradeStateDone) == 2);
    EXPECT(R, countState(s.NodeStates[UpgradeStateUpgradeRequired], UpgradeStateCordonRequired) == 1);
    EXPECT(R, countState(s.NodeStates[UpgradeStateUpgradeRequired], UpgradeStateUpgradeRequired) == 2);
  });
  R.it("maxParallel 3 and maxUnavailable 2 [:514-544]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateUpgradeRequired] = upgradeRequired(e, 5);
    DriverUpgradePolicySpec p = AUTO;
    p.MaxParallelUpgrades = 3;
    p.MaxUnavailable = IntOrString::FromInt(2);
    EXPECT(R, !apply(e, &s, &p).has_value());
    EXPECT(R, countState(s.NodeStates[UpgradeStateUpgradeRequired], UpgradeStateUpgradeRequired) == 3);
    EXPECT(R, countState(s.NodeStates[UpgradeStateUpgradeRequired], UpgradeStateCordonRequired) == 2);
  });
  auto threeIn = [&](Env& e, ClusterUpgradeState& s, const char* state) {
    for (int i = 0; i < 3; i++) s.NodeStates[state].push_back(e.entry(e.nodeWithUpgradeState(state)));
  };
  R.it("should skip pod deletion if no filter is provided [:615-633]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    threeIn(e, s, UpgradeStateWaitForJobsRequired);
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, countState(s.NodeStates[UpgradeStateWaitForJobsRequired], UpgradeStateDrainRequired) == 3);
  });
  R.it("should not skip pod deletion if a filter is provided [:634-657]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    threeIn(e, s, UpgradeStateWaitForJobsRequired);
    e.m->WithPodDeletionEnabled([](const Pod&) { return false; });
    EXPECT(R, e.m->IsPodDeletionEnabled());
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, countState(s.NodeStates[UpgradeStateWaitForJobsRequired], UpgradeStatePodDeletionRequired) == 3);
  });
  R.it("should not attempt to delete pods if pod deletion is disabled [:658-695]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    threeIn(e, s, UpgradeStatePodDeletionRequired);
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, e.pods.evictionCalls == 0);
    EXPECT(R, countState(s.NodeStates[UpgradeStatePodDeletionRequired], UpgradeStateDrainRequired) == 3);
  });
  R.it("should skip drain if it's disabled by policy [:696-729]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    threeIn(e, s, UpgradeStateDrainRequired);
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, countState(s.NodeStates[UpgradeStateDrainRequired], UpgradeStatePodRestartRequired) == 3);
    ClusterUpgradeState s2;
    threeIn(e, s2, UpgradeStateDrainRequired);
    DriverUpgradePolicySpec p = AUTO;
    p.DrainSpec = DrainSpec{};
    EXPECT(R, !apply(e, &s2, &p).has_value());
    EXPECT(R, countState(s2.NodeStates[UpgradeStateDrainRequired], UpgradeStatePodRestartRequired) == 3);
  });
  R.it("should schedule drain for UpgradeStateDrainRequired nodes and pass drain config [:730-763]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    threeIn(e, s, UpgradeStateDrainRequired);
    DriverUpgradePolicySpec p = AUTO;
    p.DrainSpec = DrainSpec{};
    p.DrainSpec->Enable = true;
    size_t seen = 0;
    const DrainSpec* spec = nullptr;
    e.drain.fn = [&](const DrainConfiguration& c) { seen = c.Nodes.size(); spec = c.Spec; return Error(); };
    EXPECT(R, !apply(e, &s, &p).has_value());
    EXPECT(R, seen == 3 && spec == &*p.DrainSpec);
    EXPECT(R, countState(s.NodeStates[UpgradeStateDrainRequired], UpgradeStateDrainRequired) == 3);
  });
  R.it("should fail if drain manager returns an error [:764-788]", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    threeIn(e, s, UpgradeStateDrainRequired);
    DriverUpgradePolicySpec p = AUTO;
    p.DrainSpec = DrainSpec{};
    p.DrainSpec->Enable = true;
    e.drain.fn = [](const DrainConfiguration&) { return Errorf("drain failed"); };
    EXPECT(R, apply(e, &s, &p).has_value());
  });
  R.it("should not restart pod if it's up to date or already terminating [:789-849]", [&] {
    Env e; fresh(e);
    Pod *up = e.pod(H, "Running"), *outRunning = e.pod(OUT, "Running"), *outTerminating = e.pod(OUT, "", {}, {}, true);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStatePodRestartRequired] = {e.entry(e.nodeWithUpgradeState(UpgradeStatePodRestartRequired), up, &e.daemonSet),
                                                    e.entry(e.nodeWithUpgradeState(UpgradeStatePodRestartRequired), outRunning, &e.daemonSet),
                                                    e.entry(e.nodeWithUpgradeState(UpgradeStatePodRestartRequired), outTerminating, &e.daemonSet)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, e.pods.restarted.size() == 1 && e.pods.restarted[0] == outRunning);
  });
  R.it("should unblock loading of the driver instead of restarting the Pod [:850-883]", [&] {
    Env e; fresh(e);
    Node* n = e.nodeWithUpgradeState(UpgradeStatePodRestartRequired);
    n->Annotations[GetUpgradeDriverWaitForSafeLoadAnnotationKey()] = "true";
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStatePodRestartRequired] = {e.entry(n, e.pod(H, "Running"), &e.daemonSet)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, n->Annotations.count(GetUpgradeDriverWaitForSafeLoadAnnotationKey()) == 0);
    EXPECT(R, e.pods.restarted.empty());
  });
  R.it("should move pod to UncordonRequired if PodRestart/UpgradeFailed, up to date and ready [:884-919]", [&] {
    Env e; fresh(e);
    Pod* ready = e.pod(H, "Running", {{true, 0}});
    Node *a = e.nodeWithUpgradeState(UpgradeStatePodRestartRequired), *b = e.nodeWithUpgradeState(UpgradeStateFailed);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStatePodRestartRequired] = {e.entry(a, ready, &e.daemonSet)};
    s.NodeStates[UpgradeStateFailed] = {e.entry(b, ready, &e.daemonSet)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(a) == UpgradeStateUncordonRequired && getNodeUpgradeState(b) == UpgradeStateUncordonRequired);
  });
  R.it("... and to UpgradeDone when the node was initially Unschedulable [:920-971]", [&] {
    Env e; fresh(e);
    Pod* ready = e.pod(H, "Running", {{true, 0}});
    Node *a = e.nodeWithUpgradeState(UpgradeStatePodRestartRequired), *b = e.nodeWithUpgradeState(UpgradeStateFailed);
    for (Node* n : {a, b}) { n->Unschedulable = true; n->Annotations[GetUpgradeInitialStateAnnotationKey()] = "true"; }
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStatePodRestartRequired] = {e.entry(a, ready, &e.daemonSet)};
    s.NodeStates[UpgradeStateFailed] = {e.entry(b, ready, &e.daemonSet)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(a) == UpgradeStateDone && getNodeUpgradeState(b) == UpgradeStateDone);
    EXPECT(R, !isUnschedulableAnnotationPresent(a) && !isUnschedulableAnnotationPresent(b));
  });
  R.it("should move pod to UpgradeFailed if the driver pod is failing with repeated restarts [:972-1017]", [&] {
    Env e; fresh(e);
    Pod* p1 = e.pod(H, "Running", {{false, 0}});
    Pod* p2 = e.pod(H, "Running", {{false, 0}}, {{true, 0}});
    Pod* p3 = e.pod(H, "Running", {{false, 11}}, {{true, 0}});
    Pod* p4 = e.pod(H, "Running", {{false, 0}}, {{false, 11}});
    Node* n[4];
    ClusterUpgradeState s;
    Pod* ps[4] = {p1, p2, p3, p4};
    for (int i = 0; i < 4; i++) {
      n[i] = e.nodeWithUpgradeState(UpgradeStatePodRestartRequired);
      s.NodeStates[UpgradeStatePodRestartRequired].push_back(e.entry(n[i], ps[i], &e.daemonSet));
    }
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(n[0]) == UpgradeStatePodRestartRequired && getNodeUpgradeState(n[1]) == UpgradeStatePodRestartRequired);
    EXPECT(R, getNodeUpgradeState(n[2]) == UpgradeStateFailed && getNodeUpgradeState(n[3]) == UpgradeStateFailed);
  });
  R.it("should move pod to ValidationRequired when validation is enabled [:1018-1053]", [&] {
    Env e; fresh(e);
    Node* n = e.nodeWithUpgradeState(UpgradeStatePodRestartRequired);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStatePodRestartRequired] = {e.entry(n, e.pod(H, "Running", {{true, 0}}), &e.daemonSet)};
    e.m->WithValidationEnabled("app=validator");
    EXPECT(R, e.m->IsValidationEnabled());
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(n) == UpgradeStateValidationRequired);
  });
  R.it("validation-required: done -> uncordon-required, not done -> unchanged [:1054-1088]", [&] {
    Env e; fresh(e);
    e.m->WithValidationEnabled("app=validator");
    Node* n = e.nodeWithUpgradeState(UpgradeStateValidationRequired);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateValidationRequired] = {e.entry(n, e.pod("", "", {}, {}, false, false), &e.daemonSet)};
    e.validation.done = false;
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(n) == UpgradeStateValidationRequired);
    e.validation.done = true;
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(n) == UpgradeStateUncordonRequired);
  });
  R.it("validation done and node initially Unschedulable -> upgrade-done, annotation removed [:1089-1127]", [&] {
    Env e; fresh(e);
    e.m->WithValidationEnabled("app=validator");
    Node* n = e.nodeWithUpgradeState(UpgradeStateValidationRequired);
    n->Unschedulable = true;
    n->Annotations[GetUpgradeInitialStateAnnotationKey()] = "true";
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateValidationRequired] = {e.entry(n, e.pod("", "", {}, {}, false, false), &e.daemonSet)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(n) == UpgradeStateDone && !isUnschedulableAnnotationPresent(n));
  });
  R.it("should uncordon UncordonRequired pod and finish upgrade [:1128-1153]", [&] {
    Env e; fresh(e);
    Node* n = e.nodeWithUpgradeState(UpgradeStateUncordonRequired);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateUncordonRequired] = {e.entry(n)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, e.cordon.uncordoned.size() == 1 && e.cordon.uncordoned[0] == n);
    EXPECT(R, getNodeUpgradeState(n) == UpgradeStateDone);
  });
  R.it("should fail if cordonManager fails [:1154-1178]", [&] {
    Env e; fresh(e);
    Node* n = e.nodeWithUpgradeState(UpgradeStateUncordonRequired);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateUncordonRequired] = {e.entry(n)};
    e.cordon.uncordon = [](Node*) { return Errorf("cordonManagerFailed"); };
    EXPECT(R, apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(n) != UpgradeStateDone);
  });
  R.it("orphaned pod: not moved to UpgradeRequired [:1180-1199]", [&] {
    Env e; fresh(e);
    Pod* orphan = e.pod("", "", {}, {}, false, false);
    Node *a = e.nodeWithUpgradeState(""), *b = e.nodeWithUpgradeState(UpgradeStateDone);
    ClusterUpgradeState s;
    s.NodeStates[""] = {e.entry(a, orphan, nullptr)};
    s.NodeStates[UpgradeStateDone] = {e.entry(b, orphan, nullptr)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(a) == UpgradeStateDone && getNodeUpgradeState(b) == UpgradeStateDone);
  });
  R.it("orphaned pod + upgrade-requested -> UpgradeRequired [:1200-1221]", [&] {
    Env e; fresh(e);
    Pod* orphan = e.pod("", "", {}, {}, false, false);
    Node *a = e.nodeWithUpgradeState(""), *b = e.nodeWithUpgradeState(UpgradeStateDone);
    a->Annotations[GetUpgradeRequestedAnnotationKey()] = b->Annotations[GetUpgradeRequestedAnnotationKey()] = "true";
    ClusterUpgradeState s;
    s.NodeStates[""] = {e.entry(a, orphan, nullptr)};
    s.NodeStates[UpgradeStateDone] = {e.entry(b, orphan, nullptr)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(a) == UpgradeStateUpgradeRequired && getNodeUpgradeState(b) == UpgradeStateUpgradeRequired);
  });
  R.it("upgrade-required orphan -> CordonRequired and upgrade-requested annotation removed [:1222-1237]", [&] {
    Env e; fresh(e);
    Node* n = e.nodeWithUpgradeState(UpgradeStateUpgradeRequired);
    n->Annotations[GetUpgradeRequestedAnnotationKey()] = "true";
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateUpgradeRequired] = {e.entry(n, e.pod("", "", {}, {}, false, false), nullptr)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(n) == UpgradeStateCordonRequired);
    EXPECT(R, n->Annotations.count(GetUpgradeRequestedAnnotationKey()) == 0);
  });
  R.it("should restart pod if it is Orphaned [:1238-1267]", [&] {
    Env e; fresh(e);
    Pod* orphan = e.pod(OUT, "Running");
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStatePodRestartRequired] = {e.entry(e.nodeWithUpgradeState(UpgradeStatePodRestartRequired), orphan, nullptr)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, e.pods.restarted.size() == 1 && e.pods.restarted[0] == orphan);
  });
  R.it("upgrade-failed with a pod lacking the hash label stays failed [:1268-1294]", [&] {
    Env e; fresh(e);
    Node* n = e.nodeWithUpgradeState(UpgradeStateFailed);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateFailed] = {e.entry(n, e.pod("", "Running", {{true, 0}}, {}, false, false), &e.daemonSet)};
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    EXPECT(R, getNodeUpgradeState(n) == UpgradeStateFailed);
  });
  // ---- requestor mode (upgrade_requestor.go) ----
  StateOptions RQ;
  RQ.Requestor.UseMaintenanceOperator = true;
  R.it("requestor: upgrade-required -> node-maintenance-required + requestor-mode annotation [:1296-1352]", [&] {
    Env e; fresh(e, RQ);
    ClusterUpgradeState s;
    for (int i = 0; i < 3; i++)
      s.NodeStates[UpgradeStateUpgradeRequired].push_back(e.entry(e.nodeWithUpgradeState(UpgradeStateUpgradeRequired), e.pod("", "Running", {{true, 0}}, {}, false, false), &e.daemonSet));
    DriverUpgradePolicySpec p = AUTO;
    p.DrainSpec = DrainSpec{};
    p.DrainSpec->Enable = true;
    EXPECT(R, !apply(e, &s, &p).has_value());
    for (auto* en : s.NodeStates[UpgradeStateUpgradeRequired]) {
      EXPECT(R, en->Node->Annotations[GetUpgradeRequestorModeAnnotationKey()] == "true");
      EXPECT(R, getNodeUpgradeState(en->Node) == UpgradeStateNodeMaintenanceRequired);
    }
  });
  R.it("requestor: NodeMaintenance Ready -> pod-restart-required; missing -> upgrade-required [:1435-1510]", [&] {
    Env e; fresh(e, RQ);
    NodeMaintenance ready;
    ready.ReadyConditionWithReasonReady = true;
    NodeMaintenance pending;
    Node *a = e.nodeWithUpgradeState(UpgradeStateNodeMaintenanceRequired), *b = e.nodeWithUpgradeState(UpgradeStateNodeMaintenanceRequired),
         *c = e.nodeWithUpgradeState(UpgradeStateNodeMaintenanceRequired);
    ClusterUpgradeState s;
    Pod* pd = e.pod("", "Running", {{true, 0}}, {}, false, false);
    s.NodeStates[UpgradeStateNodeMaintenanceRequired] = {e.entry(a, pd, &e.daemonSet, &ready), e.entry(b, pd, &e.daemonSet, nullptr), e.entry(c, pd, &e.daemonSet, &pending)};
    DriverUpgradePolicySpec p = AUTO;
    EXPECT(R, !apply(e, &s, &p).has_value());
    EXPECT(R, getNodeUpgradeState(a) == UpgradeStatePodRestartRequired);
    EXPECT(R, getNodeUpgradeState(b) == UpgradeStateUpgradeRequired);
    EXPECT(R, getNodeUpgradeState(c) == UpgradeStateNodeMaintenanceRequired);
  });
  R.it("requestor: uncordon-required with requestor-mode annotation -> upgrade-done, annotation removed [:1532-1564]", [&] {
    Env e; fresh(e, RQ);
    ClusterUpgradeState s;
    for (int i = 0; i < 3; i++) {
      Node* n = e.nodeWithUpgradeState(UpgradeStateUncordonRequired);
      n->Annotations[GetUpgradeRequestorModeAnnotationKey()] = "true";
      s.NodeStates[UpgradeStateUncordonRequired].push_back(e.entry(n, e.pod("", "Running", {{true, 0}}, {}, false, false), &e.daemonSet, &e.nm));
    }
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    for (auto* en : s.NodeStates[UpgradeStateUncordonRequired]) {
      EXPECT(R, getNodeUpgradeState(en->Node) == UpgradeStateDone);
      EXPECT(R, en->Node->Annotations.count(GetUpgradeRequestorModeAnnotationKey()) == 0);
    }
    EXPECT(R, e.cordon.uncordoned.empty());
  });
  R.it("requestor (shared): validation-required + initial-state + requestor-mode -> uncordon-required [:1566-1609]", [&] {
    Env e; fresh(e, RQ);
    ClusterUpgradeState s;
    for (int i = 0; i < 3; i++) {
      Node* n = e.nodeWithUpgradeState(UpgradeStateValidationRequired);
      n->Annotations[GetUpgradeRequestorModeAnnotationKey()] = "true";
      n->Annotations[GetUpgradeInitialStateAnnotationKey()] = "true";
      s.NodeStates[UpgradeStateValidationRequired].push_back(e.entry(n, e.pod("", "Running", {{true, 0}}, {}, false, false), &e.daemonSet, &e.nm));
    }
    EXPECT(R, !apply(e, &s, &AUTO).has_value());
    for (auto* en : s.NodeStates[UpgradeStateValidationRequired]) {
      EXPECT(R, getNodeUpgradeState(en->Node) == UpgradeStateUncordonRequired);
      EXPECT(R, en->Node->Annotations.count(GetUpgradeRequestorModeAnnotationKey()) == 1);
      EXPECT(R, !isUnschedulableAnnotationPresent(en->Node));
    }
  });
  // ---- counters (common_manager.go:715-788) on a mixed snapshot ----
  R.it("CommonUpgradeStateManager counters follow common_manager.go:715-788", [&] {
    Env e; fresh(e);
    ClusterUpgradeState s;
    s.NodeStates[UpgradeStateUpgradeRequired] = upgradeRequired(e, 4, 1);
    threeIn(e, s, UpgradeStateCordonRequired);
    threeIn(e, s, UpgradeStateDone);
    s.NodeStates[UpgradeStateFailed].push_back(e.entry(e.nodeWithUpgradeState(UpgradeStateFailed)));
    s.NodeStates[UpgradeStateNodeMaintenanceRequired].push_back(e.entry(e.nodeWithUpgradeState(UpgradeStateNodeMaintenanceRequired)));
    EXPECT(R, e.m->GetTotalManagedNodes(s) == 11);
    EXPECT(R, e.m->GetUpgradesInProgress(s) == 4);
    EXPECT(R, e.m->GetUpgradesDone(s) == 3 && e.m->GetUpgradesFailed(s) == 1 && e.m->GetUpgradesPending(s) == 4);
    EXPECT(R, e.m->GetCurrentUnavailableNodes(s) == 1);
    EXPECT(R, e.m->GetUpgradesAvailable(s, 6, 11) == 2);
    EXPECT(R, e.m->GetUpgradesAvailable(s, 0, 5) == 1);
  });
}

}  // namespace spec
