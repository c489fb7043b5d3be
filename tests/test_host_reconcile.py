"""CPU tier: the host-side mirror of pkg/reconcile (host/reconcile.hpp) applies reconcileHandler's switch
(reference reconcile.go:70-90) to status words.  Compiled as a tiny C++ program (no GPU involved)."""
import subprocess
import tempfile
import textwrap
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent

SRC = textwrap.dedent(r'''
    #include <cassert>
    #include <cstdio>
    #include "aws-global-accelerator-controller_b200/host/reconcile.hpp"
    using namespace garecon;
    struct Q : RateLimitingQueue {
      std::vector<std::string> log;
      bool Get(std::string*, bool*) override { return false; }
      void Done(const std::string& k) override { log.push_back("done:" + k); }
      void Forget(const std::string& k) override { log.push_back("forget:" + k); }
      void AddRateLimited(const std::string& k) override { log.push_back("ratelimited:" + k); }
      void AddAfter(const std::string& k, std::chrono::nanoseconds d) override {
        log.push_back("after" + std::to_string(std::chrono::duration_cast<std::chrono::seconds>(d).count()) + ":" + k);
      }
    };
    static std::string act(uint32_t status) {
      Q q;
      auto re = ResultFromStatus(status);
      ApplyResult(q, "ns/n", re.first, re.second);
      std::string s;
      for (auto& l : q.log) s += l + ";";
      return s;
    }
    int main() {
      // Result{}, nil -> Forget (reconcile.go:86-88)
      assert(act(GAR_STATUS(GAR_ST_OK, 0, 0)) == "forget:ns/n;");
      assert(act(GAR_STATUS(GAR_ST_SKIP_NO_LB, 0, 0)) == "forget:ns/n;");
      // RequeueAfter > 0 -> Forget + AddAfter (reconcile.go:79-82); 30 s LB not active, 60 s accelerator lookup
      assert(act(GAR_STATUS(GAR_ST_REQUEUE_30S, 0, 0)) == "forget:ns/n;after30:ns/n;");
      assert(act(GAR_STATUS(GAR_ST_REQUEUE_60S, GAR_D_ACCEL_NONE, 0)) == "forget:ns/n;after60:ns/n;");
      // error, not NoRetry -> AddRateLimited (reconcile.go:75-77)
      assert(act(GAR_STATUS(GAR_ST_ERR_RETRY, GAR_D_LB_NOT_FOUND, 0)) == "ratelimited:ns/n;");
      // NoRetryError -> dropped, nothing requeued (reconcile.go:73-74; pkg/errors/errors_test.go:17-35 IsNoRetry)
      assert(act(GAR_STATUS(GAR_ST_ERR_NORETRY, 0, 0)) == "");
      assert(IsNoRetry(Error::NoRetry("x")) && !IsNoRetry(Error::Retry("x")) && !IsNoRetry(Error::None()));
      puts("ok");
      return 0;
    }
''')


def test_reconcile_handler_mirror():
    with tempfile.TemporaryDirectory() as d:
        src = Path(d) / "t.cpp"
        src.write_text(SRC)
        subprocess.run(["g++", "-std=c++17", "-I", str(REPO), "-o", f"{d}/t", str(src)], check=True)
        out = subprocess.run([f"{d}/t"], capture_output=True, text=True, check=True)
        assert out.stdout.strip() == "ok"
