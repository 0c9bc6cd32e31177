// Test doubles with the semantics of the reference suite's mocks (pkg/upgrade/upgrade_suit_test.go:114-182).
#pragma once
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "../../k8s-operator-libs_b200/host/upgrade.hpp"

namespace mocks {
using namespace upgrade;

struct NodeUpgradeStateProviderMock : NodeUpgradeStateProvider {
  std::map<std::string, Node*> nodes;
  Error GetNode(const std::string& name, Node** out) override {  // upgrade_suit_test.go:131-139
    auto it = nodes.find(name);
    if (it == nodes.end()) return Errorf("node not found");
    *out = it->second;
    return std::nullopt;
  }
  Error ChangeNodeUpgradeState(Node* node, const std::string& s) override {  // :116-120
    node->Labels[GetUpgradeStateLabelKey()] = s;
    return std::nullopt;
  }
  Error ChangeNodeUpgradeAnnotation(Node* node, const std::string& k, const std::string& v) override {  // :121-130
    if (v == "null") node->Annotations.erase(k); else node->Annotations[k] = v;
    return std::nullopt;
  }
};
struct CordonManagerMock : CordonManager {
  std::function<Error(Node*)> uncordon;
  std::vector<Node*> cordoned, uncordoned;
  Error Cordon(Node* n) override { cordoned.push_back(n); return std::nullopt; }
  Error Uncordon(Node* n) override { uncordoned.push_back(n); return uncordon ? uncordon(n) : std::nullopt; }
};
struct DrainManagerMock : DrainManager {
  std::function<Error(const DrainConfiguration&)> fn;
  int calls = 0;
  Error ScheduleNodesDrain(const DrainConfiguration& c) override { calls++; return fn ? fn(c) : std::nullopt; }
};
struct PodManagerMock : PodManager {
  std::function<Error(const std::vector<Pod*>&)> restart;
  std::function<Error(const PodManagerConfig&)> eviction;
  int evictionCalls = 0, waitCalls = 0;
  std::vector<Pod*> restarted;
  Error ScheduleCheckOnPodCompletion(const PodManagerConfig&) override { waitCalls++; return std::nullopt; }
  Error SchedulePodsRestart(const std::vector<Pod*>& pods) override { restarted = pods; return restart ? restart(pods) : std::nullopt; }
  Error SchedulePodEviction(const PodManagerConfig& c) override { evictionCalls++; return eviction ? eviction(c) : std::nullopt; }
  PodDeletionFilter GetPodDeletionFilter() override { return nullptr; }
  Error GetPodControllerRevisionHash(const Pod* pod, std::string* hash) override {  // :159-168 — never errors
    auto it = pod->Labels.find(PodControllerRevisionHashLabelKey);
    *hash = it == pod->Labels.end() ? "" : it->second;
    return std::nullopt;
  }
  Error GetDaemonsetControllerRevisionHash(const DaemonSet*, std::string* hash) override {  // :169-171
    *hash = "test-hash-12345";
    return std::nullopt;
  }
};
struct ValidationManagerMock : ValidationManager {
  bool done = true;  // :179-182
  Error Validate(Node*, bool* out) override { *out = done; return std::nullopt; }
};
// the reference wires the REAL SafeDriverLoadManager over the mocked provider (safe_driver_load_manager.go:51-71)
struct SafeDriverLoadManagerImpl : SafeDriverLoadManager {
  NodeUpgradeStateProvider* provider;
  explicit SafeDriverLoadManagerImpl(NodeUpgradeStateProvider* p) : provider(p) {}
  Error IsWaitingForSafeDriverLoad(const Node* node, bool* waiting) override {
    auto it = node->Annotations.find(GetUpgradeDriverWaitForSafeLoadAnnotationKey());
    *waiting = it != node->Annotations.end() && !it->second.empty();
    return std::nullopt;
  }
  Error UnblockLoading(Node* node) override {
    bool w = false;
    IsWaitingForSafeDriverLoad(node, &w);
    if (!w) return std::nullopt;
    return provider->ChangeNodeUpgradeAnnotation(node, GetUpgradeDriverWaitForSafeLoadAnnotationKey(), "null");
  }
};
struct K8sClientMock : K8sClient {
  std::vector<DaemonSet*> daemonSets;
  std::vector<Pod*> pods;
  Error ListDaemonSets(const std::string&, const StringMap&, std::vector<DaemonSet*>* out) override { *out = daemonSets; return std::nullopt; }
  Error ListPods(const std::string&, const StringMap&, std::vector<Pod*>* out) override { *out = pods; return std::nullopt; }
};

// ---- tiny spec runner ---------------------------------------------------------------------------------------
struct Runner {
  int passed = 0, failed = 0;
  const char* current = "";
  void check(bool ok, const char* expr, const char* file, int line) {
    if (ok) return;
    failed_here = true;
    std::printf("    FAILED %s (%s:%d)\n", expr, file, line);
  }
  bool failed_here = false;
  void it(const char* name, const std::function<void()>& body) {
    current = name;
    failed_here = false;
    body();
    std::printf("%s %s\n", failed_here ? "not ok" : "ok", name);
    failed_here ? failed++ : passed++;
  }
};
#define EXPECT(r, cond) (r).check((cond), #cond, __FILE__, __LINE__)
}  // namespace mocks
