// gsbd: v1.Pod JSON -> the gsb_pod table gsb_allocate reads (podutils.go:37-131, podmanager.go:101-262)
// Private to gsbd.cc (one translation unit): everything lives in an unnamed namespace.
#ifndef GSBD_PODS_HPP_
#define GSBD_PODS_HPP_
#include <errno.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <cmath>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "gsbd_log.hpp"
#include "json.hpp"

namespace {

struct PodRec {
  std::string name, ns, uid;
  uint64_t rv = 0;  // metadata.resourceVersion when it is a decimal number (etcd's are), else 0 = "cannot compare"
};
// The pending-pod table gsb_allocate reads. Rows keep the order in which the apiserver listed (then streamed)
// them; the strings a gsb_pod points at live in heap records that never move, so an upsert touches one row.
struct PodTable {
  std::vector<std::unique_ptr<PodRec>> recs;
  std::vector<gsb_pod> pods;
  std::unordered_map<std::string, size_t> by_uid;  // live rows only
  uint64_t list_rv = 0; // resourceVersion of the LIST the table was last rebuilt from
  size_t dead = 0;      // rows deleted by a watch event: on_node = 0 makes gsb_allocate skip them entirely
  bool unique = true;   // no two live rows share a uid (always true for a table the informer maintains)
  std::chrono::steady_clock::time_point stamp;
  bool valid = false;

  void clear() {
    recs.clear();
    pods.clear();
    by_uid.clear();
    dead = 0;
    unique = true;
    list_rv = 0;
  }
  void point(size_t i) {
    pods[i].name = recs[i]->name.c_str();
    pods[i].ns = recs[i]->ns.c_str();
    pods[i].uid = recs[i]->uid.c_str();
  }
  void append(PodRec &&r, const gsb_pod &g) {
    recs.emplace_back(new PodRec(std::move(r)));
    pods.push_back(g);
    point(recs.size() - 1);
    if (!by_uid.emplace(recs.back()->uid, recs.size() - 1).second) unique = false;  // a LIST that repeats a uid
  }
  void upsert(PodRec &&r, const gsb_pod &g) {
    auto it = by_uid.find(r.uid);
    if (it == by_uid.end()) return append(std::move(r), g);
    *recs[it->second] = std::move(r);
    pods[it->second] = g;
    point(it->second);
  }
  void remove(const std::string &uid) {
    auto it = by_uid.find(uid);
    if (it == by_uid.end()) return;
    pods[it->second].on_node = 0;
    by_uid.erase(it);
    if (++dead > 64 && dead * 4 > recs.size()) compact();
  }
  void compact() {  // drop the tombstones, keeping the order of the live rows
    size_t w = 0;
    for (size_t i = 0; i < recs.size(); i++) {
      auto it = by_uid.find(recs[i]->uid);
      if (it == by_uid.end() || it->second != i) continue;
      if (w != i) {
        recs[w] = std::move(recs[i]);
        pods[w] = pods[i];
        it->second = w;
      }
      w++;
    }
    recs.resize(w);
    pods.resize(w);
    dead = 0;
  }
};

bool atoi_strict(const std::string &s, long long *out) {  // strconv.Atoi
  size_t i = (s.size() && (s[0] == '+' || s[0] == '-')) ? 1 : 0;
  if (i >= s.size()) return false;
  for (size_t k = i; k < s.size(); k++)
    if (s[k] < '0' || s[k] > '9') return false;
  errno = 0;
  const long long v = strtoll(s.c_str(), nullptr, 10);
  if (errno) return false;
  *out = v;
  return true;
}
bool parse_uint64(const std::string &s, uint64_t *out) {  // strconv.ParseUint(s, 10, 64)
  if (s.empty()) return false;
  for (char c : s)
    if (c < '0' || c > '9') return false;
  errno = 0;
  const unsigned long long v = strtoull(s.c_str(), nullptr, 10);
  if (errno) return false;
  *out = v;
  return true;
}

// one v1.Pod JSON object -> (PodRec, gsb_pod) with the fields the reference reads (podutils.go:37-131)
void pod_row(const json::Value &p, const std::string &node, PodRec *r, gsb_pod *g) {
  const json::Value *md = p.get("metadata");
  if (md) {
    if (auto *v = md->get("name")) r->name = v->str();
    if (auto *v = md->get("namespace")) r->ns = v->str();
    if (auto *v = md->get("uid")) r->uid = v->str();
    if (auto *v = md->get("resourceVersion")) parse_uint64(v->str(), &r->rv);
  }
  memset(g, 0, sizeof *g);
  g->gpu_idx = -1;
  if (const json::Value *cs = p.path({"spec", "containers"}))  // podutils.go:122-131: spec.containers only
    for (const json::Value &c : cs->arr)
      if (const json::Value *lim = c.path({"resources", "limits", kResourceName})) g->gpu_mem_limit += quantity_value(*lim);
  const json::Value *ann = md ? md->get("annotations") : nullptr;
  if (ann && ann->type == json::Value::Object) {
    if (auto *v = ann->get(kEnvResourceIndex)) {
      long long id;
      if (atoi_strict(v->str(), &id) && id >= -2147483648LL && id <= 2147483647LL) g->gpu_idx = id < 0 ? -1 : (int32_t)id;
    }
    if (auto *v = ann->get(kEnvResourceAssumeTime)) {
      g->has_assume_time = 1;
      uint64_t at;
      if (parse_uint64(v->str(), &at)) g->assume_time = at;
    }
    if (auto *v = ann->get(kEnvAssignedFlag)) {
      g->has_assigned = 1;
      g->assigned_is_false = v->str() == "false";
    }
  }
  const json::Value *nn = p.path({"spec", "nodeName"});
  g->on_node = nn && nn->str() == node;
}

// v1.PodList JSON -> table; `pending_only` = the kubelet path's phase filter (podmanager.go:101-123)
void build_table(const json::Value &list, const std::string &node, bool pending_only, PodTable *t) {
  t->clear();
  const json::Value *items = list.get("items");
  if (!items || items->type != json::Value::Array) return;
  t->recs.reserve(items->arr.size());
  t->pods.reserve(items->arr.size());
  t->by_uid.reserve(items->arr.size() * 2);
  for (const json::Value &p : items->arr) {
    if (pending_only) {
      const json::Value *ph = p.path({"status", "phase"});
      if (!ph || ph->str() != "Pending") continue;
    }
    PodRec r;
    gsb_pod g;
    pod_row(p, node, &r, &g);
    t->append(std::move(r), g);
  }
  if (const json::Value *v = list.path({"metadata", "resourceVersion"})) parse_uint64(v->str(), &t->list_rv);
}


}  // namespace

#endif  // GSBD_PODS_HPP_
