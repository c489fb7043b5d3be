// reconcile.hpp — host-side mirror of the reference's worker surface, batch edition (header-only C++17).
//
// The reference (Go) pumps one key at a time:
//     pkg/reconcile/reconcile.go:17-20   type Result struct { Requeue bool; RequeueAfter time.Duration }
//     pkg/reconcile/reconcile.go:26-42   ProcessNextWorkItem(queue, keyToObj, processDelete, processCreateOrUpdate) bool
//     pkg/reconcile/reconcile.go:44-91   reconcileHandler: (Result, error) -> Forget / AddRateLimited / AddAfter
// There is no Go toolchain in the build image, so the host side above the C ABI is written in C++ with the same
// names, argument meaning and error behaviour; INTEGRATION.md shows the cgo file a maintainer would add instead.
//
// ProcessBatch drains every key currently in the queue, runs ONE gar_diff over the snapshot the caller packed
// (informer cache + listed AWS state) and then applies, per key, exactly the switch of reconcileHandler to
// the status word the engine produced for that key's object row.  Executing the ops (the AWS SDK calls of
// global_accelerator.go:654-1013 / route53.go:183-315) stays with the caller: ExecuteOps is a callback.
#pragma once

#include <chrono>
#include <cstdint>
#include <functional>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/garecon.h"

namespace garecon {

// reconcile.Result (reconcile.go:17-20)
struct Result {
  bool Requeue = false;
  std::chrono::nanoseconds RequeueAfter{0};
};

// pkg/errors.NoRetryError / IsNoRetry (pkg/errors/errors.go:8-39): an error that must not be requeued
struct Error {
  bool set = false;
  bool noRetry = false;
  std::string msg;
  static Error None() { return {}; }
  static Error Retry(std::string m) { return {true, false, std::move(m)}; }
  static Error NoRetry(std::string m) { return {true, true, std::move(m)}; }
};
inline bool IsNoRetry(const Error &e) { return e.set && e.noRetry; }

// the subset of workqueue.RateLimitingInterface the handler uses (reconcile.go:26-91)
struct RateLimitingQueue {
  virtual ~RateLimitingQueue() = default;
  virtual bool Get(std::string *key, bool *shutdown) = 0;  // non-blocking in this mirror: false when empty
  virtual void Done(const std::string &key) = 0;
  virtual void Forget(const std::string &key) = 0;
  virtual void AddRateLimited(const std::string &key) = 0;
  virtual void AddAfter(const std::string &key, std::chrono::nanoseconds d) = 0;
};

// Status word -> (Result, error), the inverse of how the reference's process* functions return
// (globalaccelerator/service.go:54-126, route53/service.go:48-111; DESIGN.md "Status words").
inline std::pair<Result, Error> ResultFromStatus(uint32_t status) {
  using namespace std::chrono_literals;
  switch (GAR_STATUS_CODE(status)) {
    case GAR_ST_IGNORED:     // never enqueued by the reference; treated as a clean sync
    case GAR_ST_OK:
    case GAR_ST_SKIP_NO_LB:
      return {Result{}, Error::None()};
    case GAR_ST_REQUEUE_30S:
      return {Result{true, 30s}, Error::None()};  // global_accelerator.go:125-128
    case GAR_ST_REQUEUE_60S:
      return {Result{false, 60s}, Error::None()};  // route53.go:68-77 (retryAfter = 1 minute)
    case GAR_ST_ERR_RETRY:
      return {Result{}, Error::Retry("reconcile error, detail " + std::to_string(GAR_STATUS_DETAIL(status)))};
    case GAR_ST_ERR_NORETRY:
      return {Result{}, Error::NoRetry("reconcile error (no retry)")};
    case GAR_ST_PANIC:
      return {Result{}, Error::NoRetry("reference would panic: lbIngress hostname with fewer than two labels (provider.go:9-10)")};
  }
  return {Result{}, Error::Retry("unknown status")};
}

// The queue action of reconcileHandler (reconcile.go:70-90) for one key.
inline void ApplyResult(RateLimitingQueue &q, const std::string &key, const Result &res, const Error &err) {
  if (err.set) {
    if (!IsNoRetry(err)) q.AddRateLimited(key);  // :75-77; NoRetry errors are dropped (:73-74)
    return;
  }
  if (res.RequeueAfter.count() > 0) {  // :79-82
    q.Forget(key);
    q.AddAfter(key, res.RequeueAfter);
  } else if (res.Requeue) {  // :83-85
    q.AddRateLimited(key);
  } else {  // :86-88
    q.Forget(key);
  }
}

enum class Controller { GlobalAccelerator = GAR_CTRL_GA, Route53 = GAR_CTRL_R53 };

struct BatchStats {
  size_t keys = 0, forgotten = 0, requeued = 0, delayed = 0, dropped = 0, deleted_keys = 0;
};

// KeyToRow: "ns/name" -> object row of the packed snapshot, or -1 when the key is no longer in the cache
// (the reference's kerrors.IsNotFound branch, reconcile.go:61-62: processDelete; in the batch design the delete
// decisions arrive as the orphan sections of the change set, so such a key is forgotten unless its cleanup failed).
using KeyToRowFunc = std::function<int64_t(const std::string &)>;
// What went wrong while the ops were executed (host/executor.hpp ExecReport carries the same two maps): the reference returns
// the AWS error from process{Create,Update,Delete} and reconcileHandler requeues the key rate-limited (reconcile.go:75-77).
struct OpFailures {
  std::map<std::pair<uint8_t, uint32_t>, std::string> objects;  // (controller, object row) -> first error
  std::map<std::string, std::string> owners;                    // "service/ns/name" of a key that left the cache -> error of its cleanup
};
using ExecuteOpsFunc = std::function<OpFailures(const gar_changeset &)>;

// Batch counterpart of `for ProcessNextWorkItem(...) {}` (globalaccelerator/controller.go:222-230): one diff,
// then reconcileHandler's switch for every drained key.  `ctrl` + `kind` name the queue (the reference has a Service and an
// Ingress queue per controller); `executeOps` must execute the ops of THIS controller only (ExecuteChangeSet's filter) — the
// other controller's worker executes the other half — and report what failed: a key whose ops hit an AWS error is requeued
// rate-limited whatever its status word said.
inline int ProcessBatch(gar_engine *engine, RateLimitingQueue &queue, Controller ctrl, uint8_t kind, KeyToRowFunc keyToRow, ExecuteOpsFunc executeOps,
                        BatchStats *stats = nullptr, std::string *error = nullptr) {
  std::vector<std::string> keys;
  for (;;) {
    std::string k;
    bool shutdown = false;
    if (!queue.Get(&k, &shutdown) || shutdown) break;
    keys.push_back(std::move(k));
  }
  BatchStats st;
  st.keys = keys.size();
  gar_changeset cs{};
  int rc = gar_diff(engine, &cs);
  if (rc != GAR_OK) {
    // no CPU fallback: the whole batch is a retryable error (SURVEY.md §8.4 row b)
    if (error) *error = gar_last_error(engine);
    for (auto &k : keys) {
      queue.AddRateLimited(k);
      queue.Done(k);
      st.requeued++;
    }
    if (stats) *stats = st;
    return rc;
  }
  OpFailures failures;
  if (executeOps) failures = executeOps(cs);
  const uint32_t *status = ctrl == Controller::GlobalAccelerator ? cs.status_ga : cs.status_r53;
  const std::string resource = kind == GAR_KIND_SERVICE ? "service/" : "ingress/";
  for (auto &k : keys) {
    int64_t row = keyToRow(k);
    if (row < 0 || row >= (int64_t)cs.n_objects) {
      st.deleted_keys++;
      if (failures.owners.count(resource + k)) {  // processDelete returned the AWS error (service.go:41-44)
        queue.AddRateLimited(k);
        st.requeued++;
      } else {
        queue.Forget(k);  // processDelete returned (Result{}, nil): the orphan ops carried the cleanup
        st.forgotten++;
      }
    } else if (failures.objects.count({(uint8_t)ctrl, (uint32_t)row})) {
      ApplyResult(queue, k, Result{}, Error::Retry(failures.objects[{(uint8_t)ctrl, (uint32_t)row}]));
      st.requeued++;
    } else {
      auto re = ResultFromStatus(status[row]);
      ApplyResult(queue, k, re.first, re.second);
      if (re.second.set) (IsNoRetry(re.second) ? st.dropped : st.requeued)++;
      else if (re.first.RequeueAfter.count() > 0) st.delayed++;
      else if (re.first.Requeue) st.requeued++;
      else st.forgotten++;
    }
    queue.Done(k);
  }
  gar_changeset_free(engine, &cs);
  if (stats) *stats = st;
  return GAR_OK;
}

}  // namespace garecon
