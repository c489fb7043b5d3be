// BASELINE config 1 ("reconcile 100 synthetic Service type:LoadBalancer ... 1 worker") through the host-side mirror
// of the reference's worker surface (host/reconcile.hpp, mirror of pkg/reconcile/reconcile.go:26-91) and the C ABI.
// The in-memory snapshot stands in for the mock cloudprovider the reference does not have.
// Prints one JSON line with the queue actions; tests/test_gpu_worker.py checks it against the status words.
#include <cstdio>
#include <deque>
#include <map>
#include <string>

#include "aws-global-accelerator-controller_b200/host/reconcile.hpp"
#include "aws-global-accelerator-controller_b200/synth/gsyn.h"

using namespace garecon;

struct Queue : RateLimitingQueue {
  std::deque<std::string> pending;
  int done = 0, forget = 0, ratelimited = 0, after30 = 0, after60 = 0;
  bool Get(std::string *key, bool *shutdown) override {
    *shutdown = false;
    if (pending.empty()) return false;
    *key = pending.front();
    pending.pop_front();
    return true;
  }
  void Done(const std::string &) override { done++; }
  void Forget(const std::string &) override { forget++; }
  void AddRateLimited(const std::string &) override { ratelimited++; }
  void AddAfter(const std::string &, std::chrono::nanoseconds d) override {
    (std::chrono::duration_cast<std::chrono::seconds>(d).count() == 30 ? after30 : after60)++;
  }
};

int main(int argc, char **argv) {
  int cfgid = argc > 1 ? atoi(argv[1]) : 1;
  uint32_t n = argc > 2 ? (uint32_t)atoi(argv[2]) : 100;
  gsyn_config cfg;
  gsyn_preset(cfgid, n, &cfg);
  gsyn_snapshot *snap = gsyn_generate(&cfg);
  const gar_objects *o = gsyn_objects(snap);
  gar_config ec{GAR_ABI_VERSION, 0, cfg.cluster, 0};
  gar_engine *e = nullptr;
  if (gar_engine_create(&ec, &e) != GAR_OK) {
    fprintf(stderr, "engine: %s\n", gar_last_error(nullptr));
    return 2;
  }
  if (gar_snapshot_load(e, o, gsyn_actual(snap)) != GAR_OK) {
    fprintf(stderr, "load: %s\n", gar_last_error(e));
    return 3;
  }
  // serviceQueue of the globalaccelerator controller: every Service key, plus one key that left the cache
  std::map<std::string, int64_t> rowOf;
  Queue q;
  for (uint32_t i = 0; i < o->n_objects; i++) {
    if (o->obj_kind[i] != GAR_KIND_SERVICE) continue;
    std::string key((const char *)o->slab + GAR_STR_OFF(o->obj_ns[i]), GAR_STR_LEN(o->obj_ns[i]) + 1 + GAR_STR_LEN(o->obj_name[i]));
    rowOf[key] = i;
    q.pending.push_back(key);
  }
  q.pending.push_back("default/deleted-service");
  uint64_t n_ops = 0;
  BatchStats st;
  std::string err;
  int rc = ProcessBatch(
      e, q, Controller::GlobalAccelerator, GAR_KIND_SERVICE,
      [&](const std::string &k) -> int64_t {
        auto it = rowOf.find(k);
        return it == rowOf.end() ? -1 : it->second;
      },
      [&](const gar_changeset &cs) {
        n_ops = cs.n_ops;
        return OpFailures{};
      },
      &st, &err);
  printf("{\"rc\": %d, \"keys\": %zu, \"forgotten\": %zu, \"requeued\": %zu, \"delayed\": %zu, \"dropped\": %zu, \"deleted_keys\": %zu, \"n_ops\": %llu, "
         "\"q_done\": %d, \"q_forget\": %d, \"q_ratelimited\": %d, \"q_after30\": %d, \"q_after60\": %d}\n",
         rc, st.keys, st.forgotten, st.requeued, st.delayed, st.dropped, st.deleted_keys, (unsigned long long)n_ops, q.done, q.forget, q.ratelimited, q.after30, q.after60);
  gar_engine_destroy(e);
  gsyn_free(snap);
  return rc == GAR_OK ? 0 : 1;
}
